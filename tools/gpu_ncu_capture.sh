# ncu launch list + --set full capture of one bf16 forward (B=256); summarise with tools/ncu_summarize.py.
mkdir -p gpurun_out/r2k
cd /root/repo
K='regex:conv_tc_kernel|enc0_kernel|gemm_tc_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 25 -c 25 --csv --log-file gpurun_out/r2k/launches_bf16_b256.csv python tools/one_forward.py 256 2 > gpurun_out/r2k/launches.log 2>&1; echo "launches rc $?" >> gpurun_out/r2k/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 25 -c 25 -o gpurun_out/r2k/prof_r02 python tools/one_forward.py 256 2 > gpurun_out/r2k/full.log 2>&1; echo "full rc $?" >> gpurun_out/r2k/rc.txt
ls -la gpurun_out/r2k; cat gpurun_out/r2k/rc.txt
