for lib in rs_ray_toy_b200/variants/librrt_old.so rs_ray_toy_b200/librrt_sm100.so; do
  for cfg in "c5 --scale 0.5 --nsamp 129" "c2 --nsamp 9"; do
    echo "== $lib $cfg"
    RRT_LIB=$lib python tools/bench_render.py --config $cfg --reps 3 2>&1 | grep -o '"rep": [0-9]*\|"Msamples_per_s": [0-9.]*\|mean rgb.*' | paste - - | tail -3
  done
done
