set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2e_pytest.log; tail -6 gpurun_out/r2e_pytest.log
python bench.py --steps 3 --warmup 3 --no-strong > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; cut -c1-200 gpurun_out/r2e_bench_n1.json; tail -3 gpurun_out/r2e_bench_n1.err
