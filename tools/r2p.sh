mkdir -p gpurun_out/r2p
cd /root/repo
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2p/pytest.txt 2>&1; echo "pytest rc $?" >> gpurun_out/r2p/rc.txt
cat gpurun_out/r2p/rc.txt; tail -n 5 gpurun_out/r2p/pytest.txt
