# GPU box: full GPU suite, the N=1 bench record, launch lists at the product's chunk size.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/gputests.txt 2>&1; tail -3 gpurun_out/gputests.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json
for c in c5 c4; do
  python tools/bench_render.py --config $c --reps 1 > gpurun_out/render_$c.txt 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_render_${c}_full.csv \
      python tools/bench_render.py --config $c --reps 1 > gpurun_out/ncu_$c.log 2>&1
  tail -2 gpurun_out/render_$c.txt
done
