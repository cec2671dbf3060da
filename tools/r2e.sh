mkdir -p gpurun_out/r2e
cd /root/repo
timeout 60 python tools/tn_dbg.py 0 16 32 48 4 52 0 > gpurun_out/r2e/tn_dbg.txt 2>&1; echo "dbg rc $?" >> gpurun_out/r2e/rc.txt
WUNET_TN_NA=4 timeout 60 python tools/tn_dbg.py 0 >> gpurun_out/r2e/tn_dbg.txt 2>&1
timeout 400 python -m pytest tests/test_train_gpu.py -q -s > gpurun_out/r2e/pytest_train.txt 2>&1; echo "train tests rc $?" >> gpurun_out/r2e/rc.txt
timeout 120 python tools/sp_time.py > gpurun_out/r2e/sp_time.txt 2>&1; echo "sp_time rc $?" >> gpurun_out/r2e/rc.txt
timeout 400 python -m pytest tests/test_parity_gpu.py -q -s -k "fp32_tc" > gpurun_out/r2e/pytest_sp.txt 2>&1; echo "sp tests rc $?" >> gpurun_out/r2e/rc.txt
timeout 300 python bench.py --mode train --steps 5 --warmup 2 > gpurun_out/r2e/bench_train.json 2> gpurun_out/r2e/bench_train.err; echo "bench train rc $?" >> gpurun_out/r2e/rc.txt
cat gpurun_out/r2e/rc.txt gpurun_out/r2e/tn_dbg.txt gpurun_out/r2e/sp_time.txt
tail -n 6 gpurun_out/r2e/pytest_train.txt gpurun_out/r2e/pytest_sp.txt
