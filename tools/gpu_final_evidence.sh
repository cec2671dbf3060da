# The round's final evidence pass on one GPU: ncu launch list, ncu --set full capture, bench.py, GPU tests.
mkdir -p gpurun_out/r2z
cd /root/repo
O=gpurun_out/r2z
K='regex:conv_tc_kernel|enc0_kernel|gemm_tc_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 25 -c 25 --csv --log-file $O/launches_bf16_b256.csv python tools/one_forward.py 256 2 > $O/launches.log 2>&1; echo "launches rc $?" >> $O/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 25 -c 25 -o $O/prof_r02 python tools/one_forward.py 256 2 > $O/full.log 2>&1; echo "full rc $?" >> $O/rc.txt
timeout 600 python bench.py > $O/bench_bf16_b256.json 2> $O/bench_bf16_b256.err; echo "fwd rc $?" >> $O/rc.txt
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc $?" >> $O/rc.txt
ls -la $O; cat $O/rc.txt; tail -n 3 $O/pytest_gpu.txt; cut -c1-300 $O/bench_bf16_b256.json
