# gpurun -- 'bash tools/gpu_pair_ab2.sh': the last block's row-pair form with single-phase producers against the default path
cd /root/repo
O=gpurun_out/pair3
mkdir -p $O gpurun_out/pair
rm -f gpurun_out/pair/results.jsonl
run() { timeout -k 5 150 python tools/pair_check.py "$1" "$2" > $O/log_$3.txt 2>&1; echo "pair_check $1 $3 rc $?" >> $O/rc.txt; tail -c 300 $O/log_$3.txt | tail -n 1; }
run 1 "" default
run 3 "" pair3
run 3 "24:mt=1,small=1,na=2" pair3_na2
cp gpurun_out/pair/results.jsonl $O/results.jsonl; rm -f gpurun_out/pair/blk_ref.npz
cat $O/rc.txt
