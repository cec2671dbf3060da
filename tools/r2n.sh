mkdir -p gpurun_out/r2n
cd /root/repo
timeout 120 python tools/sp_time.py > gpurun_out/r2n/sp_time.txt 2>&1; echo "sp_time rc $?" >> gpurun_out/r2n/rc.txt
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py tests/test_enhance.py -m gpu -q > gpurun_out/r2n/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2n/rc.txt
cat gpurun_out/r2n/rc.txt; tail -n 3 gpurun_out/r2n/pytest.txt; cat gpurun_out/r2n/sp_time.txt
