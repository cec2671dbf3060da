# gpurun -- 'bash tools/gpu_pair_sweep.sh': block 0 on the tensor cores (WUNET_TC_ENC0), tilings of blocks 0 / 1 in their row-group
# forms, and the GPU tests with the modes on
cd /root/repo
O=gpurun_out/pair2
mkdir -p $O gpurun_out/pair
rm -f gpurun_out/pair/results.jsonl
run() { timeout -k 5 150 python tools/pair_check.py "$@" > $O/log_$3.txt 2>&1; echo "pair_check $1 $3 rc $?" >> $O/rc.txt; tail -c 400 $O/log_$3.txt | tail -n 1; }
run 0 "" default
WUNET_TC_ENC0=1 run 0 "" enc0tc
WUNET_TC_ENC0=1 timeout -k 5 150 python tools/ovr_try.py "0:mt=2" "0:small=1" "0:mt=1,na=3" "0:mt=1,nacc=1" > $O/sweep_blk0.txt 2>&1; echo "sweep0 rc $?" >> $O/rc.txt
WUNET_TC_PAIR=1 timeout -k 5 200 python tools/ovr_try.py "1:mt=1,na=2" "1:mt=1,na=3" "1:mt=1,na=4" "1:mt=2,na=3" "1:mt=4,na=2,res=0" "1:mt=2,na=2,res=0" "1:mt=1,small=1,na=2" "1:mt=1,small=1,na=3" "1:mt=2,small=1,na=2" > $O/sweep_blk1.txt 2>&1; echo "sweep1 rc $?" >> $O/rc.txt
WUNET_TC_PAIR=1 timeout -k 5 300 python -m pytest tests -m gpu -q > $O/pytest_pair1.txt 2>&1; echo "pytest pair1 rc $?" >> $O/rc.txt
WUNET_TC_ENC0=1 timeout -k 5 300 python -m pytest tests -m gpu -q > $O/pytest_enc0tc.txt 2>&1; echo "pytest enc0tc rc $?" >> $O/rc.txt
cp gpurun_out/pair/results.jsonl $O/results.jsonl; rm -f gpurun_out/pair/blk_ref.npz
cat $O/rc.txt; grep -v "wunet tc\|wunet gemm" $O/sweep_blk0.txt | tail -n 6; grep -v "wunet tc\|wunet gemm" $O/sweep_blk1.txt | tail -n 11; tail -n 4 $O/pytest_pair1.txt; tail -n 12 $O/pytest_enc0tc.txt
