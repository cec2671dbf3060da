mkdir -p gpurun_out/r2f
cd /root/repo
timeout 600 python bench.py > gpurun_out/r2f/bench_bf16_b256.json 2> gpurun_out/r2f/bench.err; echo "fwd rc $?" >> gpurun_out/r2f/rc.txt
cat gpurun_out/r2f/rc.txt; cut -c1-200 gpurun_out/r2f/bench_bf16_b256.json
