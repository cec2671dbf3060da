mkdir -p gpurun_out/r2r
cd /root/repo
WUNET_LIB_PATH=$PWD/wave_u_net_for_speech_enhancement_b200/build/libwunet_b200_trace.so timeout 120 python tools/trace_levels.py 23 22 24 1 2> gpurun_out/r2r/trace.txt > gpurun_out/r2r/trace.out; echo "trace rc $?" >> gpurun_out/r2r/rc.txt
timeout 150 python tools/ovr_try.py "1:res=0" "1:res=0,bulk=1" "2:res=0" "2:res=0,bulk=1" "3:bulk=1" "3:mt=4" "1:mt=2" "2:mt=4" "5:bulk=1" "6:bulk=1" "21:bulk=1" "19:bulk=1" > gpurun_out/r2r/ovr.txt 2>&1; echo "ovr rc $?" >> gpurun_out/r2r/rc.txt
cat gpurun_out/r2r/rc.txt; cat gpurun_out/r2r/ovr.txt; wc -c gpurun_out/r2r/trace.txt
