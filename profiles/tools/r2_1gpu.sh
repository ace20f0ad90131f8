#!/bin/bash
# One-GPU round-2 sweep: the GPU test suite, the default bench line (headline + the other BASELINE configurations), BASELINE
# config 5 at its full size on ONE GPU (strong-scaling base), config 1 at its native size, and the Laplacian-kernel A/B on
# graphs without 16-bit indices.
#   gpurun --timeout 2400 -- 'profiles/tools/r2_1gpu.sh r2'
tag=${1:-r2}
show() { python - "$@" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[2], 'ms_per_step', d['ms_per_step'], 'value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for w in d.get('workloads') or []:
    print('   ', {k: w.get(k) for k in ('workload', 'ms_per_step', 'value', 'frac_algorithmic', 'frac_hoisted', 'frac_traffic', 'kernels_per_step', 'error')})
PY
}
(timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25) > gpurun_out/${tag}_gputests.log 2>&1; tail -3 gpurun_out/${tag}_gputests.log
python bench.py --steps 20 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; show gpurun_out/${tag}_bench_n1.json default
python bench.py --workload wave3d --size 693 --total --steps 20 --no-cpu --no-workloads > gpurun_out/${tag}_wave693_n1.json 2> gpurun_out/${tag}_wave693_n1.err; show gpurun_out/${tag}_wave693_n1.json wave693
python bench.py --size 100 --steps 200 --no-weighted --no-workloads > gpurun_out/${tag}_heat100_n1.json 2>/dev/null; show gpurun_out/${tag}_heat100_n1.json heat100
for w in wave3d cml; do
  GDS_B200_LAP_KERNEL=lapT python bench.py --workload $w --steps 20 --no-cpu --no-workloads > gpurun_out/${tag}_${w}_lapT.json 2>/dev/null; show gpurun_out/${tag}_${w}_lapT.json "$w lapT"
done
