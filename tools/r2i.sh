mkdir -p gpurun_out/r2i
cd /root/repo
timeout 150 python tools/ab_check.py WUNET_TC_PDL > gpurun_out/r2i/pdl_ab.txt 2>&1; echo "ab rc $?" >> gpurun_out/r2i/rc.txt
cat gpurun_out/r2i/rc.txt; tail -n 8 gpurun_out/r2i/pdl_ab.txt
