mkdir -p gpurun_out/r2h
cd /root/repo
WUNET_TC_DEBUG=1 timeout 150 python tools/ab_check.py WUNET_TC_GEMM > gpurun_out/r2h/gemm_ab.txt 2>&1; echo "ab rc $?" >> gpurun_out/r2h/rc.txt
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py tests/test_enhance.py -m gpu -q -s > gpurun_out/r2h/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2h/rc.txt
cat gpurun_out/r2h/rc.txt; grep -v "wunet tc\]" gpurun_out/r2h/gemm_ab.txt | tail -n 14; tail -n 5 gpurun_out/r2h/pytest.txt
