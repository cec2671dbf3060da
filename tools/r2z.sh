mkdir -p gpurun_out/r2z
cd /root/repo
L=/root/repo/wave_u_net_for_speech_enhancement_b200/build
for pr in 256 128 64 0 256; do
  WUNET_TC_L2PROMO=$pr WUNET_LIB_PATH=$L/libw_promo.so timeout 120 python tools/lib_times.py 256 bf16 >> gpurun_out/r2z/times.txt 2>&1; echo "promo $pr rc $?" >> gpurun_out/r2z/rc.txt
done
cat gpurun_out/r2z/rc.txt; cat gpurun_out/r2z/times.txt
