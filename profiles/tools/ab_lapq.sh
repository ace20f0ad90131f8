#!/bin/bash
# Same-box A/B of the quad Laplacian kernel (heat RK4, grid 7072^2): environment switches of the default build and
# alternative builds under gds_b200/lib/ab/ (built with `nvcc ... -DLAPQ_MINB=.. -DLAPQ_UNROLL=..`, see Makefile EXTRA).
# One JSON line per configuration in gpurun_out/ab_<name>.json; summary on stdout.
B="python bench.py --steps 20 --no-cpu --no-weighted"
run() { name=$1; shift; env "$@" $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open(f'gpurun_out/ab_{name}.json') if l.startswith('{')][-1])
    print(f"{name:10s} ms_per_step={d['ms_per_step']:.4f} launch_ms={d['roofline']['launch_ms']:.4f} kernel={d['roofline']['kernel']}")
except Exception as e:
    print(name, 'failed', e)
PY
}
run default A=1
run nouniw GDS_B200_NO_UNIW=1
run runs0 GDS_B200_LAPQ_RUNS=0
for v in "$@"; do
  [ -f gds_b200/lib/ab/$v.so ] && run $v GDS_B200_LIB=$PWD/gds_b200/lib/ab/$v.so
done
run default2 A=1
