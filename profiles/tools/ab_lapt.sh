#!/bin/bash
# Same-box A/B for the WEIGHTED heat problem (grid 7072^2, random edge weights): lapT builds under gds_b200/lib/ab/
# and the weighted quad kernel (GDS_B200_LAP_KERNEL=quad).  One JSON line per configuration in gpurun_out/abw_<name>.json.
B="python bench.py --steps 20 --no-cpu --weighted"
run() { name=$1; shift; env "$@" $B > gpurun_out/abw_$name.json 2> gpurun_out/abw_$name.err; python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open(f'gpurun_out/abw_{name}.json') if l.startswith('{')][-1])
    print(f"{name:10s} ms_per_step={d['ms_per_step']:.4f} launch_ms={d['roofline']['launch_ms']:.4f}")
except Exception as e:
    print(name, 'failed', e)
PY
}
run lapT A=1
run quad GDS_B200_LAP_KERNEL=quad
for v in "$@"; do
  [ -f gds_b200/lib/ab/$v.so ] && run $v GDS_B200_LIB=$PWD/gds_b200/lib/ab/$v.so
done
