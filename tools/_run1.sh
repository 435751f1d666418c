set -x
nproc; free -g | head -2; df -h /dev/shm /tmp | cat; lscpu | grep -i -E "numa|model name|socket" | head; nvidia-smi topo -m 2>/dev/null | head -20
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_pytest.log; tail -5 gpurun_out/r2a_pytest.log
python tools/host_probe.py > gpurun_out/r2a_host_probe.log 2>&1; head -8 gpurun_out/r2a_host_probe.log
python tools/host_probe.py 3840 2160 2 2>&1 | head -4 >> gpurun_out/r2a_host_probe.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cut -c1-600 gpurun_out/r2a_bench.json
