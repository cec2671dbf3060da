mkdir -p gpurun_out/r2v
cd /root/repo
L=/root/repo/wave_u_net_for_speech_enhancement_b200/build
WUNET_LIB_PATH=$L/libw_hd_trace.so timeout 200 python tools/trace_levels.py 24 2> gpurun_out/r2v/trace_hd.txt; echo "trace rc $?" >> gpurun_out/r2v/rc.txt
cat gpurun_out/r2v/rc.txt; grep -c role gpurun_out/r2v/trace_hd.txt
