mkdir -p gpurun_out/r2u
cd /root/repo
L=/root/repo/wave_u_net_for_speech_enhancement_b200/build
export WUNET_TC_DEBUG=1
WUNET_LIB_PATH=$L/libw_hd.so timeout 200 python tools/ovr_try.py "24:small=1,na=3" "24:small=1,na=2" "24:small=1,na=3,nacc=1" > gpurun_out/r2u/hd1.txt 2>&1; echo "hd1 rc $?" >> gpurun_out/r2u/rc.txt
cat gpurun_out/r2u/rc.txt; grep -E "blk 24|default|block 24" gpurun_out/r2u/hd1.txt
