set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_golden.py tests/test_gpu_render.py -m gpu -q -rP -k "golden or lane or mandelbrot_interior or benchmarked" 2>&1 | grep -E "4K bands|passed|failed|FAILED|Error" | tail -20 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log
# launch lists (device time per launch)
for t in c2 tiled1 mandelbrot literal-mandelbrot tetration literal-tetration raymarch literal-raymarch visualizer; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launch_$t.csv python tools/ncu_target.py $t 3 > /dev/null 2>&1
  grep -v "^==" gpurun_out/r2c_launch_$t.csv | awk -F'","' 'NR>1{print $5, $NF}' | tail -4
done
capture() {   # name, kernel regex, target, launches
  ncu --set full --clock-control none -k regex:$2 -s 1 -c 1 -o /tmp/$1 -f python tools/ncu_target.py $3 $4 > /dev/null 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r2c_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/r2c_$1_details.csv 2>/dev/null
}
capture rows_c2 visualizer_rows_kernel c2 3
capture final_c2 final_kernel c2 3
capture tiled1 visualizer_tiled_kernel tiled1 3
capture stft stft_mel_kernel stft 3
capture rows_4k visualizer_rows_kernel visualizer 3
for t in mandelbrot tetration raymarch; do
  capture $t frame_lanes_kernel $t 2
  capture literal_$t frame_kernel literal-$t 2
done
du -sh gpurun_out; ls gpurun_out | head -50
