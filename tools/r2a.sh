mkdir -p gpurun_out/r2a
cd /root/repo
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
WUNET_TEST_NATIVE_TRAIN=1 timeout 400 python -m pytest tests/test_train_gpu.py -q > gpurun_out/r2a/train.txt 2>&1; echo "train rc $?" >> gpurun_out/r2a/rc.txt
WUNET_TEST_BF16_MODEL=1 timeout 300 python -m pytest tests/test_bf16_model_gpu.py -q > gpurun_out/r2a/bf16model.txt 2>&1; echo "bf16model rc $?" >> gpurun_out/r2a/rc.txt
for m in 1 2 4 3 7; do timeout 200 python tools/exp_check.py $m > gpurun_out/r2a/exp_$m.txt 2>&1; echo "exp $m rc $?" >> gpurun_out/r2a/rc.txt; done
timeout 300 python tools/incumbent.py 1 64 256 > gpurun_out/r2a/incumbent.txt 2>&1; echo "incumbent rc $?" >> gpurun_out/r2a/rc.txt
cat gpurun_out/r2a/rc.txt
tail -5 gpurun_out/r2a/train.txt gpurun_out/r2a/bf16model.txt
