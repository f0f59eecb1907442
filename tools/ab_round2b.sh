# GPU box: deposit warp-sum A/B, chunk-size variants at full size, launch lists and one ncu --set full of the
# per-kind shade kernel.  Every ncu pass follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
run() {  # lib cfg...
  lib=$1; shift 1
  echo "== $lib $*"
  RRT_LIB=$lib python tools/bench_render.py --config $* 2>&1 | grep -o '"rep": [0-9]*\|"Msamples_per_s": [0-9.]*\|"launches": [0-9]*\|mean rgb.*' | paste - - - | tail -3
}
P=rs_ray_toy_b200/librrt_sm100.so
for cfg in "c4 --scale 0.5 --reps 3" "c5 --scale 0.5 --nsamp 129 --reps 3"; do
  run $P $cfg
  run rs_ray_toy_b200/variants/librrt_dep0.so $cfg
done
for cfg in "c4 --reps 3" "c5 --reps 2"; do
  run $P $cfg
  for v in rs_ray_toy_b200/variants/librrt_chunk*.so; do run $v $cfg; done
done
python -m pytest tests/test_gpu_render.py -x -q -m gpu 2>&1 | tail -2
for c in c4 c5; do
  python tools/bench_render.py --config $c --scale 0.5 --reps 1 > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_render_$c.csv \
      python tools/bench_render.py --config $c --scale 0.5 --reps 1 > gpurun_out/ncu_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:shade_range_kernel -s 3 -c 2 -o gpurun_out/shade_range_c4 \
    python tools/bench_render.py --config c4 --scale 0.5 --reps 1 > gpurun_out/ncu_full_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shade_range_kernel -s 2 -c 2 -o gpurun_out/shade_range_c5 \
    python tools/bench_render.py --config c5 --scale 0.5 --nsamp 129 --reps 1 > gpurun_out/ncu_full_c5.log 2>&1
ls -la gpurun_out
