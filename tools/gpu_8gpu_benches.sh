# gpurun --gpus 8 -- 'bash tools/gpu_8gpu_benches.sh': train / enhance / forward benches on 8 GPUs and the 2-rank NCCL test.
mkdir -p gpurun_out/r2m
cd /root/repo
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --mode train --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2m/bench_train_8gpu.json 2> gpurun_out/r2m/bench_train_8gpu.err; echo "train8 rc $?" >> gpurun_out/r2m/rc.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --mode enhance --gpus 8 --steps 20 > gpurun_out/r2m/bench_enh_8gpu.json 2> gpurun_out/r2m/bench_enh_8gpu.err; echo "enh8 rc $?" >> gpurun_out/r2m/rc.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 50 --warmup 5 --no-incumbent > gpurun_out/r2m/bench_fwd_8gpu.json 2> gpurun_out/r2m/bench_fwd_8gpu.err; echo "fwd8 rc $?" >> gpurun_out/r2m/rc.txt
timeout 200 python -m pytest tests/test_ddp_gpu.py -m gpu -x -q > gpurun_out/r2m/pytest_ddp.txt 2>&1; echo "ddp rc $?" >> gpurun_out/r2m/rc.txt
cat gpurun_out/r2m/rc.txt; cut -c1-1200 gpurun_out/r2m/bench_train_8gpu.json; cut -c1-900 gpurun_out/r2m/bench_enh_8gpu.json; cut -c1-600 gpurun_out/r2m/bench_fwd_8gpu.json; tail -3 gpurun_out/r2m/pytest_ddp.txt
