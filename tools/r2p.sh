mkdir -p gpurun_out/r2p
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2p/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2p/rc.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p/smoke.txt 2>&1; echo "smoke rc $?" >> gpurun_out/r2p/rc.txt
timeout 300 python bench.py --mode train --steps 10 --warmup 2 > gpurun_out/r2p/bench_train.json 2> gpurun_out/r2p/bench_train.err; echo "train rc $?" >> gpurun_out/r2p/rc.txt
cat gpurun_out/r2p/rc.txt; tail -n 5 gpurun_out/r2p/pytest.txt; cat gpurun_out/r2p/smoke.txt | tail -4; python -c "
import json; d=json.load(open('gpurun_out/r2p/bench_train.json')); print(d['value'], d['ms_per_step'], d['incumbent'])"
