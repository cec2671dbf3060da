mkdir -p gpurun_out/r2j
cd /root/repo
export NCCL_DEBUG=WARN
timeout 300 python -m pytest tests/test_ddp_gpu.py -m gpu -q -s > gpurun_out/r2j/pytest_ddp.txt 2>&1; echo "ddp test rc $?" >> gpurun_out/r2j/rc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --mode train --gpus 2 --steps 10 --warmup 2 > gpurun_out/r2j/bench_train_2gpu.json 2> gpurun_out/r2j/bench_train_2gpu.err; echo "train2 rc $?" >> gpurun_out/r2j/rc.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --mode enhance --gpus 2 --steps 10 > gpurun_out/r2j/bench_enh_2gpu.json 2> gpurun_out/r2j/bench_enh_2gpu.err; echo "enh2 rc $?" >> gpurun_out/r2j/rc.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2j/bench_fwd_2gpu.json 2> gpurun_out/r2j/bench_fwd_2gpu.err; echo "fwd2 rc $?" >> gpurun_out/r2j/rc.txt
cat gpurun_out/r2j/rc.txt; tail -n 4 gpurun_out/r2j/pytest_ddp.txt; cat gpurun_out/r2j/bench_train_2gpu.json | cut -c1-1500
