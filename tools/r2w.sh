mkdir -p gpurun_out/r2w
cd /root/repo
L=/root/repo/wave_u_net_for_speech_enhancement_b200/build
for so in fence nofence fence nofence; do
  WUNET_LIB_PATH=$L/libw_$so.so timeout 120 python tools/lib_times.py 256 bf16 >> gpurun_out/r2w/times.txt 2>&1; echo "$so rc $?" >> gpurun_out/r2w/rc.txt
done
for so in fence nofence; do
  WUNET_LIB_PATH=$L/libw_$so.so timeout 120 python tools/lib_times.py 256 fp32_tc >> gpurun_out/r2w/times.txt 2>&1; echo "$so tc rc $?" >> gpurun_out/r2w/rc.txt
  WUNET_LIB_PATH=$L/libw_$so.so timeout 120 python tools/lib_times.py 64 bf16 >> gpurun_out/r2w/times.txt 2>&1; echo "$so b64 rc $?" >> gpurun_out/r2w/rc.txt
done
WUNET_LIB_PATH=$L/libw_nofence.so timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py -m gpu -x -q > gpurun_out/r2w/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2w/rc.txt
cat gpurun_out/r2w/rc.txt; cat gpurun_out/r2w/times.txt; tail -n 5 gpurun_out/r2w/pytest.txt
