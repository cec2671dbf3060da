mkdir -p gpurun_out/r2d
cd /root/repo
timeout 60 python tools/tn_dbg.py 0 16 32 48 4 52 0 > gpurun_out/r2d/tn_dbg2.txt 2>&1; echo "dbg rc $?" >> gpurun_out/r2d/rc.txt
WUNET_TN_NA=4 timeout 60 python tools/tn_dbg.py 0 > gpurun_out/r2d/tn_dbg3.txt 2>&1
cat gpurun_out/r2d/tn_dbg2.txt gpurun_out/r2d/tn_dbg3.txt
