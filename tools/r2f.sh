mkdir -p gpurun_out/r2f
cd /root/repo
timeout 60 python tools/tn_dbg.py 0 4 16 32 1 2 0 > gpurun_out/r2f/tn_dbg.txt 2>&1; echo "dbg rc $?" >> gpurun_out/r2f/rc.txt
timeout 120 python tools/ab_check.py WUNET_TC_TN > gpurun_out/r2f/tn_ab.txt 2>&1; echo "ab rc $?" >> gpurun_out/r2f/rc.txt
timeout 400 python -m pytest tests/test_train_gpu.py -q -s > gpurun_out/r2f/pytest_train.txt 2>&1; echo "train tests rc $?" >> gpurun_out/r2f/rc.txt
WUNET_TC_TN=1 timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py -q -s -k "bf16" > gpurun_out/r2f/pytest_tn.txt 2>&1; echo "tn tests rc $?" >> gpurun_out/r2f/rc.txt
cat gpurun_out/r2f/rc.txt gpurun_out/r2f/tn_dbg.txt gpurun_out/r2f/tn_ab.txt
grep -n "full architecture\|tuned vs naive\|passed\|failed" gpurun_out/r2f/pytest_train.txt gpurun_out/r2f/pytest_tn.txt
