mkdir -p gpurun_out/r2y
cd /root/repo
O=gpurun_out/r2y
timeout 600 python bench.py > $O/bench_bf16_b256.json 2> $O/bench_bf16_b256.err; echo "fwd rc $?" >> $O/rc.txt
timeout 400 python bench.py --mode train --steps 10 --warmup 3 > $O/bench_train_b64_1gpu.json 2> $O/bench_train.err; echo "train rc $?" >> $O/rc.txt
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference_cpu.json 2> $O/bench_reference.err; echo "ref rc $?" >> $O/rc.txt
K='regex:conv_tc_kernel|enc0_kernel|gemm_tc_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 25 -c 25 --csv --log-file $O/launches_bf16_b256.csv python tools/one_forward.py 256 2 > $O/launches.log 2>&1; echo "launches rc $?" >> $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 25 -c 25 -o $O/prof_r02 python tools/one_forward.py 256 2 > $O/full.log 2>&1; echo "full rc $?" >> $O/rc.txt
ls -la $O; cat $O/rc.txt; cut -c1-700 $O/bench_bf16_b256.json
