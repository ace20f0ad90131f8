#!/bin/bash
# One `ncu --set full` capture of ONE STEP of every BASELINE workload (all kernels of the step, in order), plus the
# C4 gather with and without the internal renumbering (L2 hit rate of the gathered vector).  One GPU.
#   gpurun --timeout 1500 -- 'profiles/tools/ncu_round2.sh r2'
# Each capture runs only after the same command has exited 0 without ncu.  Reports land in gpurun_out/<tag>_ncu_*.ncu-rep;
# read them with profiles/tools/ncu_summary.py (-> profiles/<tag>_ncu_*.csv, profiles/traffic.json).
tag=${1:-r2}
NCU="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-weighted --no-workloads --no-preroll"
cap() {   # name, kernels per step, extra env, bench args...
  name=$1; kps=$2; envs=$3; shift 3
  if env $envs $B "$@" > gpurun_out/${tag}_plain_${name}.json 2> gpurun_out/${tag}_plain_${name}.err; then
    env $envs $NCU -s $((kps * 2)) -c $kps -o gpurun_out/${tag}_ncu_${name} $B "$@" > gpurun_out/${tag}_ncu_${name}.log 2>&1
    echo "$name: captured (rc=$?)"
    # gpurun brings back at most 64 MiB: keep the raw metric table of every capture, the report itself only when asked
    ncu -i gpurun_out/${tag}_ncu_${name}.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${name}_raw.csv 2>/dev/null
    case " $KEEP_REPORTS " in *" $name "*) ;; *) rm -f gpurun_out/${tag}_ncu_${name}.ncu-rep ;; esac
  else
    echo "$name: plain run failed"; tail -3 gpurun_out/${tag}_plain_${name}.err
  fi
}
KEEP_REPORTS=${KEEP_REPORTS:-"cml ns_cavity"}
[ -n "$WITH_HEAT_UNIT" ] && cap heat_unit 5 A=1 --workload heat2d
# (launches captured: two steps' worth, so that the capture holds one complete step whatever it starts with)
cap heat_weighted 10 A=1 --workload heat2d --weighted
cap hex_advdiff 52 A=1 --workload hex_advdiff
cap ns_cavity 10 A=1 --workload ns_cavity
cap wave3d 4 A=1 --workload wave3d
cap cml 4 A=1 --workload cml
cap cml_norenumber 4 GDS_B200_NO_RENUMBER=1 --workload cml
cap cml_noshadow 4 GDS_B200_NO_SHADOW=1 --workload cml
ls -la gpurun_out/${tag}_ncu_*; du -sh gpurun_out
