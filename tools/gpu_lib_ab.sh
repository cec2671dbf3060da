# A/B of variant builds of the library (WUNET_SO_OUT=... WUNET_WAIT_NS=... python -m wave_u_net_for_speech_enhancement_b200.build --force)
# on a GPU box: per-block times of each build through tools/lib_times.py. Run with gpurun -- 'bash tools/gpu_lib_ab.sh'.
mkdir -p gpurun_out/r2s
cd /root/repo
L=/root/repo/wave_u_net_for_speech_enhancement_b200/build
for so in base all1000 nomma1000 all200 hd base; do
  WUNET_LIB_PATH=$L/libw_$so.so timeout 150 python tools/lib_times.py 256 bf16 >> gpurun_out/r2s/times.txt 2>&1; echo "$so rc $?" >> gpurun_out/r2s/rc.txt
done
WUNET_TC_HEADK=0 WUNET_LIB_PATH=$L/libw_hd.so timeout 150 python tools/lib_times.py 256 bf16 >> gpurun_out/r2s/times.txt 2>&1; echo "hd-off rc $?" >> gpurun_out/r2s/rc.txt
for so in base all1000; do
  WUNET_LIB_PATH=$L/libw_$so.so timeout 150 python tools/lib_times.py 256 fp32_tc >> gpurun_out/r2s/times.txt 2>&1; echo "$so tc rc $?" >> gpurun_out/r2s/rc.txt
  WUNET_LIB_PATH=$L/libw_$so.so timeout 150 python tools/lib_times.py 1 bf16 >> gpurun_out/r2s/times.txt 2>&1; echo "$so b1 rc $?" >> gpurun_out/r2s/rc.txt
done
WUNET_LIB_PATH=$L/libw_hd.so timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py -m gpu -x -q > gpurun_out/r2s/pytest_hd.txt 2>&1; echo "pytest hd rc $?" >> gpurun_out/r2s/rc.txt
cat gpurun_out/r2s/rc.txt; cat gpurun_out/r2s/times.txt; tail -n 5 gpurun_out/r2s/pytest_hd.txt
