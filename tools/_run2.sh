set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_pytest.log; tail -15 gpurun_out/r2b_pytest.log
python tools/host_probe.py 2>&1 | head -4 > gpurun_out/r2b_host_probe.log; cat gpurun_out/r2b_host_probe.log
python tools/host_probe.py 3840 2160 2 2>&1 | head -4 >> gpurun_out/r2b_host_probe.log; tail -3 gpurun_out/r2b_host_probe.log
