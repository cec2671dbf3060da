# gpurun -- 'bash tools/gpu_pair_sweep.sh': tilings of block 1 in row-pair mode, then the GPU tests with the mode on
cd /root/repo
O=gpurun_out/pair2
mkdir -p $O
WUNET_TC_PAIR=1 timeout -k 5 200 python tools/ovr_try.py "1:mt=1,na=2" "1:mt=1,na=3" "1:mt=1,na=4" "1:mt=2,na=3" "1:mt=4,na=2,res=0" "1:mt=2,na=2,res=0" "1:mt=1,small=1,na=2" "1:mt=1,small=1,na=3" "1:mt=2,small=1,na=2" > $O/sweep_blk1.txt 2>&1; echo "sweep rc $?" >> $O/rc.txt
WUNET_TC_PAIR=1 timeout -k 5 300 python -m pytest tests -m gpu -q > $O/pytest_pair1.txt 2>&1; echo "pytest pair1 rc $?" >> $O/rc.txt
cat $O/rc.txt; grep -v "wunet tc\|wunet gemm" $O/sweep_blk1.txt | tail -n 12; tail -n 5 $O/pytest_pair1.txt
