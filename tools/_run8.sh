set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_render.py tests/test_gpu_golden.py tests/test_gpu_scene.py -m gpu -q 2>&1 | tail -5
for dbg in 0 4; do
  SFB_ROWS_DEBUG=$dbg ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launch_vis_$dbg.csv python tools/ncu_target.py visualizer 4 > /dev/null 2>&1
  SFB_ROWS_DEBUG=$dbg ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launch_c2_$dbg.csv python tools/ncu_target.py c2 4 > /dev/null 2>&1
done
grep -h -E "rows_kernel|final_kernel" gpurun_out/r2h_launch_*.csv | awk -F'","' '{print FILENAME, $5, $NF}' | cut -c1-160
