#!/bin/bash
# Same-box A/B of the term-list kernels (configs 2 and 3): alternative builds under gds_b200/lib/ab/ made with
#   make LIB=gds_b200/lib/ab/<name>.so EXTRA="-DET_MINB=.. -DTK_MINB=.."
# One JSON line per (build, workload) in gpurun_out/abt_<name>_<workload>.json; summary on stdout.
#   gpurun --timeout 400 -- 'profiles/tools/ab_terms.sh et5_tk5 et6_tk6'
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
for wl in "ns_cavity 3000" "hex_advdiff 1500"; do
  set -- $wl "$@"; w=$1; s=$2; shift 2
  run default $w $s A=1
  for v in "$@"; do
    [ -f gds_b200/lib/ab/$v.so ] && run $v $w $s GDS_B200_LIB=$PWD/gds_b200/lib/ab/$v.so
  done
done
