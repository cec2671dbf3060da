mkdir -p gpurun_out/r2o
cd /root/repo
timeout 90 python tools/ab_check.py WUNET_TC_PFLATE > gpurun_out/r2o/pflate_ab.txt 2>&1; echo "ab rc $?" >> gpurun_out/r2o/rc.txt
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py -m gpu -q > gpurun_out/r2o/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2o/rc.txt
cat gpurun_out/r2o/rc.txt; tail -n 3 gpurun_out/r2o/pytest.txt; tail -n 8 gpurun_out/r2o/pflate_ab.txt
