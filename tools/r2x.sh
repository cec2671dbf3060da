mkdir -p gpurun_out/r2x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2x/pytest_gpu.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2x/rc.txt
timeout 300 python bench.py --mode enhance --steps 20 > gpurun_out/r2x/bench_enhance.json 2> gpurun_out/r2x/bench_enhance.err; echo "enh rc $?" >> gpurun_out/r2x/rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x/smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/r2x/rc.txt
cat gpurun_out/r2x/rc.txt; tail -n 6 gpurun_out/r2x/pytest_gpu.txt; cut -c1-1500 gpurun_out/r2x/bench_enhance.json; tail -n 3 gpurun_out/r2x/smoke.txt
