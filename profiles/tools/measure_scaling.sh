#!/bin/bash
# Weak scaling of the heat problem on N GPUs of one box, plus the bit-identity check (profiles/README.md section 3).
#   gpurun --gpus 8 --timeout 900 -- 'profiles/tools/measure_scaling.sh r2 8'
tag=${1:-rX}; nmax=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 20 --no-cpu --no-weighted > gpurun_out/${tag}_scale_heat2d_n1.json 2>/dev/null
for n in 2 4 8; do
  [ $n -gt $nmax ] && break
  $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 20 \
      > gpurun_out/${tag}_scale_heat2d_n$n.json 2> gpurun_out/${tag}_scale_heat2d_n$n.err
done
python - "$tag" <<'PY'
import glob, json, sys
tag = sys.argv[1]
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f'gpurun_out/{tag}_scale_heat2d_n{n}.json') if l.startswith('{')][-1])
    except Exception:
        continue
    base = base or d['value']
    print(f"N={n} ms_per_step={d['ms_per_step']:.4f} value={d['value']:.3e} speedup={d['value'] / base:.2f} "
          f"launch_ms_per_rank={d['roofline'].get('launch_ms_per_rank')}")
PY
n=$(( nmax < 2 ? 2 : nmax ))
timeout 240 $TR --nproc-per-node $n --master-port 29519 tests/mgpu_check.py 2>&1 | grep "bit-identical" | sort | uniq -c | sort -rn | head -40
