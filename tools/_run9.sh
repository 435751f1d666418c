set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2i_pytest.log; tail -4 gpurun_out/r2i_pytest.log
python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; cut -c1-200 gpurun_out/r2i_bench_n1.json; tail -2 gpurun_out/r2i_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2i_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2i_bench_ref.json
python tools/pcie_probe.py > gpurun_out/r2i_pcie_n1.json 2>/dev/null; cat gpurun_out/r2i_pcie_n1.json
