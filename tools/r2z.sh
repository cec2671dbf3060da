mkdir -p gpurun_out/r2z
cd /root/repo
L=/root/repo/wave_u_net_for_speech_enhancement_b200/build
for so in nohoist hoist nohoist hoist; do
  WUNET_LIB_PATH=$L/libw_$so.so timeout 120 python tools/lib_times.py 256 bf16 >> gpurun_out/r2z/times.txt 2>&1; echo "$so rc $?" >> gpurun_out/r2z/rc.txt
done
cat gpurun_out/r2z/rc.txt; cat gpurun_out/r2z/times.txt
