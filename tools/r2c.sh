mkdir -p gpurun_out/r2c
cd /root/repo
WUNET_TC_DEBUG=1 timeout 200 python tools/ab_check.py WUNET_TC_TN > gpurun_out/r2c/tn.txt 2>&1; echo "tn rc $?" >> gpurun_out/r2c/rc.txt
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2c/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2c/rc.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-incumbent --no-cpu-baseline > gpurun_out/r2c/bench_fwd.json 2> gpurun_out/r2c/bench_fwd.err; echo "bench rc $?" >> gpurun_out/r2c/rc.txt
cat gpurun_out/r2c/rc.txt
grep -v "wunet t" gpurun_out/r2c/tn.txt | tail -n 12
tail -n 5 gpurun_out/r2c/pytest.txt
