mkdir -p gpurun_out/r2b
cd /root/repo
WUNET_TC_TN=0 timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2b/pytest_tn0.txt 2>&1; echo "pytest(tn0) rc $?" >> gpurun_out/r2b/rc.txt
timeout 200 python tools/ab_check.py WUNET_TC_MERGE > gpurun_out/r2b/merge.txt 2>&1; echo "merge rc $?" >> gpurun_out/r2b/rc.txt
WUNET_TC_DEBUG=1 timeout 200 python tools/ab_check.py WUNET_TC_TN > gpurun_out/r2b/tn.txt 2>&1; echo "tn rc $?" >> gpurun_out/r2b/rc.txt
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py tests/test_enhance.py -m gpu -q -s > gpurun_out/r2b/pytest_tn1.txt 2>&1; echo "pytest(tn1) rc $?" >> gpurun_out/r2b/rc.txt
WUNET_TC_TN=0 timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r2b/bench_fwd_tn0.json 2> gpurun_out/r2b/bench_fwd_tn0.err; echo "bench(tn0) rc $?" >> gpurun_out/r2b/rc.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-incumbent --no-cpu-baseline > gpurun_out/r2b/bench_fwd_tn1.json 2> gpurun_out/r2b/bench_fwd_tn1.err; echo "bench(tn1) rc $?" >> gpurun_out/r2b/rc.txt
timeout 400 python bench.py --mode train --steps 5 --warmup 2 > gpurun_out/r2b/bench_train.json 2> gpurun_out/r2b/bench_train.err; echo "train rc $?" >> gpurun_out/r2b/rc.txt
WUNET_TC_TN=0 timeout 200 python bench.py --mode enhance --steps 10 > gpurun_out/r2b/bench_enh.json 2> gpurun_out/r2b/bench_enh.err; echo "enh rc $?" >> gpurun_out/r2b/rc.txt
cat gpurun_out/r2b/rc.txt
tail -n 8 gpurun_out/r2b/pytest_tn0.txt gpurun_out/r2b/tn.txt
