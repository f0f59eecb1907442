for c in 1048576 2097152 4194304; do for t in 0 1; do
RRT_HOST_CHUNK=$c RRT_HOST_TAPER=$t python bench.py --steps 10 --warmup 3 --no-cpu-baseline --path-config none > gpurun_out/e2e_sw.json 2>gpurun_out/e2e_sw.err
python -c "
import json; j=json.loads(open('gpurun_out/e2e_sw.json').read().strip().splitlines()[-1]); print('RRT_HOST_CHUNK=$c RRT_HOST_TAPER=$t value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1))"
done; done
