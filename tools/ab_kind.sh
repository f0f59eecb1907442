# A/B of shading by material kind (GPU box): the full GPU suite on the product build, then configs 4 and 5 at half
# resolution with the one general shade kernel (RRT_SHADE_BY_KIND=0), the per-kind kernels, and the occupancy variants.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/kind_gputests.txt 2>&1; tail -3 gpurun_out/kind_gputests.txt
run() {  # lib mode cfg...
  lib=$1; mode=$2; shift 2
  echo "== $lib by_kind=$mode $*"
  RRT_LIB=$lib RRT_SHADE_BY_KIND=$mode python tools/bench_render.py --config $* --reps 3 2>&1 | grep -o '"rep": [0-9]*\|"Msamples_per_s": [0-9.]*\|"launches": [0-9]*\|mean rgb.*' | paste - - - | tail -3
}
for cfg in "c4 --scale 0.5" "c5 --scale 0.5 --nsamp 129"; do
  run rs_ray_toy_b200/librrt_sm100.so 0 $cfg
  run rs_ray_toy_b200/librrt_sm100.so 1 $cfg
  for v in rs_ray_toy_b200/variants/librrt_k*.so; do run $v 1 $cfg; done
done
