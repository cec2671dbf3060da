# gpurun -- 'bash tools/gpu_tests.sh': the GPU test suite and the smoke check of the checked-out tree
mkdir -p gpurun_out/final
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.txt 2>&1; echo "pytest rc $?" >> gpurun_out/final/rc.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/final/rc.txt
cat gpurun_out/final/rc.txt; tail -n 4 gpurun_out/final/pytest_gpu.txt; tail -n 4 gpurun_out/final/smoke.txt
