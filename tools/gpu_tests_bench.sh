# gpurun -- 'bash tools/gpu_tests_bench.sh': GPU tests, smoke() and bench.py of the checked-out tree (no profiler)
cd /root/repo
O=gpurun_out/r2f
mkdir -p $O
timeout -k 5 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc $?" >> $O/rc.txt
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; echo "smoke rc $?" >> $O/rc.txt
timeout -k 5 600 python bench.py > $O/bench_bf16_b256.json 2> $O/bench_bf16_b256.err; echo "bench rc $?" >> $O/rc.txt
cat $O/rc.txt; tail -n 6 $O/pytest_gpu.txt; tail -n 4 $O/smoke.txt; cut -c1-400 $O/bench_bf16_b256.json
