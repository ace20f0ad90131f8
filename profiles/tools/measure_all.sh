#!/bin/bash
# One-GPU sweep of every BASELINE workload (profiles/README.md section 2) plus the pending GPU tests.
#   gpurun --timeout 900 -- 'profiles/tools/measure_all.sh r2'
# Writes gpurun_out/<tag>_bench_<workload>_n1.json (one JSON line each) and a summary on stdout.
tag=${1:-rX}
python -m pytest tests -m gpu -x -q -rxX 2>&1 | tail -15
for w in heat2d hex_advdiff ns_cavity cml wave3d; do
  extra=""; [ "$w" != heat2d ] && extra="--no-cpu"
  python bench.py --workload $w --steps 20 $extra > gpurun_out/${tag}_bench_${w}_n1.json 2> gpurun_out/${tag}_bench_${w}_n1.err
  python - "$tag" "$w" <<'PY'
import json, sys
tag, w = sys.argv[1:3]
try:
    d = json.loads([l for l in open(f'gpurun_out/{tag}_bench_{w}_n1.json') if l.startswith('{')][-1])
    r = d['roofline']
    print(f"{w:12s} ms_per_step={d['ms_per_step']:.4f} value={d['value']:.3e} frac={r['frac']:.3f} "
          f"frac_kernel_bytes={r.get('frac_kernel_bytes')} launches={d['gpu_launches']}")
except Exception as e:
    print(w, 'failed', e)
PY
done
python bench.py --size 100 --steps 200 --no-weighted > gpurun_out/${tag}_bench_heat2d_100_n1.json 2>/dev/null
