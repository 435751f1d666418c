set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2g_pytest.log; tail -6 gpurun_out/r2g_pytest.log
SFB_NO_TMA=1 python -m pytest tests/test_gpu_render.py tests/test_gpu_golden.py -m gpu -q -k "separable or benchmarked or screen_pass_into" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launch_c2.csv python tools/ncu_target.py c2 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launch_visualizer.csv python tools/ncu_target.py visualizer 3 > /dev/null 2>&1
capture() {   # name, kernel regex, target, launches
  ncu --set full --clock-control none -k regex:$2 -s 1 -c 1 -o /tmp/$1 -f python tools/ncu_target.py $3 $4 > /dev/null 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r2g_$1_raw.csv 2>/dev/null
}
capture final_c2 final_kernel c2 3
capture stft stft_mel_kernel stft 3
capture tetration frame_lanes_kernel tetration 2
capture rows_4k visualizer_rows_kernel visualizer 3
capture rows_c2 visualizer_rows_kernel c2 3
python bench.py --steps 5 --warmup 3 --no-strong > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; cut -c1-200 gpurun_out/r2g_bench_n1.json; tail -3 gpurun_out/r2g_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
