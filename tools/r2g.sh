mkdir -p gpurun_out/r2g
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2g/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2g/rc.txt
timeout 120 python tools/sp_time.py > gpurun_out/r2g/sp_time.txt 2>&1; echo "sp_time rc $?" >> gpurun_out/r2g/rc.txt
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/r2g/bench_fwd.json 2> gpurun_out/r2g/bench_fwd.err; echo "bench rc $?" >> gpurun_out/r2g/rc.txt
timeout 300 python bench.py --mode train --steps 10 --warmup 2 > gpurun_out/r2g/bench_train.json 2> gpurun_out/r2g/bench_train.err; echo "train rc $?" >> gpurun_out/r2g/rc.txt
timeout 200 python bench.py --mode enhance --steps 10 > gpurun_out/r2g/bench_enh.json 2> gpurun_out/r2g/bench_enh.err; echo "enh rc $?" >> gpurun_out/r2g/rc.txt
timeout 200 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2g/bench_ref.json 2> gpurun_out/r2g/bench_ref.err; echo "ref rc $?" >> gpurun_out/r2g/rc.txt
cat gpurun_out/r2g/rc.txt; tail -n 4 gpurun_out/r2g/pytest.txt; cat gpurun_out/r2g/sp_time.txt
