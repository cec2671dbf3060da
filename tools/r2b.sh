mkdir -p gpurun_out/r2b
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2b/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2b/rc.txt
timeout 200 python tools/merge_check.py > gpurun_out/r2b/merge.txt 2>&1; echo "merge rc $?" >> gpurun_out/r2b/rc.txt
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r2b/bench_fwd.json 2> gpurun_out/r2b/bench_fwd.err; echo "bench rc $?" >> gpurun_out/r2b/rc.txt
timeout 400 python bench.py --mode train --steps 5 --warmup 2 > gpurun_out/r2b/bench_train.json 2> gpurun_out/r2b/bench_train.err; echo "train rc $?" >> gpurun_out/r2b/rc.txt
timeout 200 python bench.py --mode enhance --steps 10 > gpurun_out/r2b/bench_enh.json 2> gpurun_out/r2b/bench_enh.err; echo "enh rc $?" >> gpurun_out/r2b/rc.txt
cat gpurun_out/r2b/rc.txt
tail -n 15 gpurun_out/r2b/pytest.txt
