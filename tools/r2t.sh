mkdir -p gpurun_out/r2t
cd /root/repo
L=/root/repo/wave_u_net_for_speech_enhancement_b200/build
for so in hd base; do
  WUNET_LIB_PATH=$L/libw_$so.so timeout 120 python tools/lib_times.py 256 bf16 >> gpurun_out/r2t/times.txt 2>&1; echo "$so rc $?" >> gpurun_out/r2t/rc.txt
done
WUNET_LIB_PATH=$L/libw_hd.so timeout 120 python tools/lib_times.py 1 bf16 >> gpurun_out/r2t/times.txt 2>&1; echo "hd b1 rc $?" >> gpurun_out/r2t/rc.txt
WUNET_LIB_PATH=$L/libw_hd.so timeout 120 python tools/lib_times.py 3 bf16 >> gpurun_out/r2t/times.txt 2>&1; echo "hd b3 rc $?" >> gpurun_out/r2t/rc.txt
WUNET_LIB_PATH=$L/libw_base.so timeout 120 python tools/lib_times.py 3 bf16 >> gpurun_out/r2t/times.txt 2>&1; echo "base b3 rc $?" >> gpurun_out/r2t/rc.txt
WUNET_LIB_PATH=$L/libw_hd.so timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py -m gpu -x -q > gpurun_out/r2t/pytest_hd.txt 2>&1; echo "pytest hd rc $?" >> gpurun_out/r2t/rc.txt
cat gpurun_out/r2t/rc.txt; cat gpurun_out/r2t/times.txt; tail -n 15 gpurun_out/r2t/pytest_hd.txt
