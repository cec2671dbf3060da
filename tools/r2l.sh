mkdir -p gpurun_out/r2l
cd /root/repo
timeout 400 python -m pytest tests/test_parity_gpu.py -q -s -k "fp32_tc" > gpurun_out/r2l/pytest_sp.txt 2>&1; echo "sp tests rc $?" >> gpurun_out/r2l/rc.txt
timeout 120 python tools/sp_time.py > gpurun_out/r2l/sp_time.txt 2>&1; echo "sp_time rc $?" >> gpurun_out/r2l/rc.txt
cat gpurun_out/r2l/rc.txt; tail -n 4 gpurun_out/r2l/pytest_sp.txt; cat gpurun_out/r2l/sp_time.txt
