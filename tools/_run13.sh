cd $GRAFT_REPO_ROOT
echo "== tests, default (ntex 0) incl. new final kernel"; python -m pytest tests/test_gpu_render.py tests/test_gpu_golden.py -m gpu -q 2>&1 | tail -3
echo "== tests with SFB_ROWS_TEX=20"; SFB_ROWS_TEX=20 python -m pytest tests/test_gpu_render.py tests/test_gpu_golden.py -m gpu -q -rP -k "separable or benchmarked or screen_pass_into or export_frame" 2>&1 | grep -E "4K bands|passed|failed|FAILED|Error|assert" | tail -12
for n in 0 4 8 12 20; do
  SFB_ROWS_TEX=$n ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_vis_$n.csv python tools/ncu_target.py visualizer 4 > /dev/null 2>&1
  SFB_ROWS_TEX=$n ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_c2_$n.csv python tools/ncu_target.py c2 4 > /dev/null 2>&1
  echo "ntex $n: 4K $(grep rows_kernel gpurun_out/r2m_vis_$n.csv | tail -1 | awk -F'","' '{print $NF}') C2 $(grep rows_kernel gpurun_out/r2m_c2_$n.csv | tail -1 | awk -F'","' '{print $NF}') final $(grep -E 'final' gpurun_out/r2m_c2_$n.csv | tail -1 | awk -F'","' '{print $NF}')"
done
