#!/bin/bash
# Same-box A/B, second sweep: carried-operand quad kernel of the coupled map lattice (LAPQ_MINB_EPI) and the face rows of
# the term-list kernel (TK_MINB_F); builds under gds_b200/lib/ab/ (make LIB=... EXTRA=...).  Summary on stdout.
run() { name=$1; wl=$2; size=$3; shift 3
  env "$@" python bench.py --steps 10 --no-cpu --no-weighted --no-workloads --no-preroll --workload $wl --size $size \
      > gpurun_out/abt_${name}_$wl.json 2> gpurun_out/abt_${name}_$wl.err
  python - "$name" "$wl" <<'PY'
import json, sys
name, wl = sys.argv[1:3]
try:
    d = json.loads([l for l in open(f'gpurun_out/abt_{name}_{wl}.json') if l.startswith('{')][-1])
    print(f"{name:10s} {wl:12s} ms_per_step={d['ms_per_step']:.4f}")
except Exception as e:
    print(name, wl, 'failed', e)
PY
}
L=$PWD/gds_b200/lib/ab
run default2 cml 4000000 A=1
run q10 cml 4000000 GDS_B200_LIB=$L/q10.so
run q12 cml 4000000 GDS_B200_LIB=$L/q12.so
run default2 ns_cavity 3000 A=1
run f6 ns_cavity 3000 GDS_B200_LIB=$L/f6.so
run f8 ns_cavity 3000 GDS_B200_LIB=$L/f8.so
