# gpurun -- 'bash tools/gpu_pair_ab.sh': A/B of the row-pair mode (WUNET_TC_PAIR) on one B200, one process per configuration
cd /root/repo
O=gpurun_out/pair
mkdir -p $O
rm -f $O/results.jsonl
run() { timeout -k 5 150 python tools/pair_check.py "$@" > $O/log_$1_${3:-x}.txt 2>&1; echo "pair_check $1 ${2:-} rc $?" >> $O/rc.txt; tail -c 600 $O/log_$1_${3:-x}.txt | tail -n 2; }
run 0
run 1
run 2
run 3
run 2 "24:mt=1,small=1,na=2" na2
timeout -k 5 240 python -m pytest tests -m gpu -x -q > $O/pytest_default.txt 2>&1; echo "pytest default rc $?" >> $O/rc.txt
WUNET_TC_PAIR=3 timeout -k 5 240 python -m pytest tests/test_parity_gpu.py tests/test_bf16_model_gpu.py tests/test_enhance.py -m gpu -q > $O/pytest_pair3.txt 2>&1; echo "pytest pair3 rc $?" >> $O/rc.txt
rm -f $O/blk_ref.npz
cat $O/rc.txt; tail -n 3 $O/pytest_default.txt; tail -n 12 $O/pytest_pair3.txt
