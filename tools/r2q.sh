mkdir -p gpurun_out/r2q
cd /root/repo
timeout 400 python -m pytest tests/test_parity_gpu.py -q -s -k "fp32_tc" > gpurun_out/r2q/pytest_sp.txt 2>&1; echo "sp tests rc $?" >> gpurun_out/r2q/rc.txt
timeout 120 python tools/sp_time.py > gpurun_out/r2q/sp_time.txt 2>&1; echo "sp_time rc $?" >> gpurun_out/r2q/rc.txt
timeout 150 python tools/ovr_try.py "23:mt=2,na=4" "23:mt=2,na=3" "23:mt=2,na=2" "23:mt=4,na=3" "23:mt=3,na=3" "22:mt=2,na=4" "22:mt=1,na=4" "22:mt=4,na=2" "24:mt=2,small=0,na=4" "24:mt=4,small=0,na=3" "24:mt=1,small=1,na=3" "24:mt=2,small=1,na=3" "21:mt=2,na=4" "21:mt=1,na=4" "20:mt=2,na=4" "20:mt=1,na=4" > gpurun_out/r2q/ovr.txt 2>&1; echo "ovr rc $?" >> gpurun_out/r2q/rc.txt
cat gpurun_out/r2q/rc.txt; tail -n 4 gpurun_out/r2q/pytest_sp.txt; cat gpurun_out/r2q/sp_time.txt; cat gpurun_out/r2q/ovr.txt
