#!/usr/bin/env python
"""Summarise `ncu --set full` reports of one step (profiles/tools/ncu_round2.sh) into committed CSV files and
profiles/traffic.json (DRAM bytes per step = sum over the step's kernels of dram__bytes_read.sum + dram__bytes_write.sum).

    python profiles/tools/ncu_summary.py r2 heat_unit:heat2d:7072 hex_advdiff:hex_advdiff:2000 ...
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']


def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    u = unit.lower()
    return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}.get(u, 1)


def to_us(val, unit):
    v = float(val.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6, 's': 1e6}.get(unit, 1)


def main():
    tag = sys.argv[1]
    traffic_path = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        traffic = json.load(open(traffic_path))
    except Exception:
        traffic = {}
    for spec in sys.argv[2:]:
        name, workload, size = spec.split(':')
        rep = os.path.join(ROOT, 'gpurun_out', f'{tag}_ncu_{name}.ncu-rep')
        raw_csv = os.path.join(ROOT, 'gpurun_out', f'{tag}_ncu_{name}_raw.csv')
        if os.path.exists(raw_csv):
            raw = open(raw_csv).read()
        elif os.path.exists(rep):
            raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        else:
            print('missing', rep)
            continue
        rows = list(csv.reader(io.StringIO(raw)))
        head, units, body = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(head)}
        # one step = the kernels up to and including a tail_kernel (Dirichlet overwrite + step counter, the last kernel of
        # every step): keep the last complete step of the capture, whatever number of launches was asked for
        names = [r[col['Kernel Name']] for r in body]
        tails = [i for i, nm in enumerate(names) if nm.startswith('tail_kernel')]
        if len(tails) >= 2:
            body = body[tails[-2] + 1:tails[-1] + 1]
        elif len(tails) == 1:     # the capture starts on a step boundary: the kernels up to and including the tail
            body = body[:tails[0] + 1]
        out_rows, total, total_us = [], 0.0, 0.0
        for r in body:
            rec = {}
            for k in KEEP:
                if k in col:
                    rec[k] = r[col[k]]
                    if k.startswith('dram__bytes'):
                        rec[k] = to_bytes(r[col[k]], units[col[k]])
                    if k == 'gpu__time_duration.sum':
                        rec[k] = to_us(r[col[k]], units[col[k]])
            total += rec.get('dram__bytes_read.sum', 0) + rec.get('dram__bytes_write.sum', 0)
            total_us += rec.get('gpu__time_duration.sum', 0)
            out_rows.append(rec)
        dst = os.path.join(ROOT, 'profiles', f'{tag}_ncu_full_{name}.csv')
        with open(dst, 'w', newline='') as fh:
            w = csv.DictWriter(fh, fieldnames=[k for k in KEEP if k in col])
            w.writeheader()
            w.writerows(out_rows)
        print(f'{name}: {len(out_rows)} kernels, DRAM bytes per step {total:.4g}, sum of kernel times {total_us:.1f} us '
              f'(cold, serialised) -> {os.path.relpath(dst, ROOT)}')
        key = size if not name.endswith(('weighted', 'norenumber', 'noshadow')) else f'{size}_{name}'
        traffic.setdefault(workload, {}).setdefault('per_step', {})[key] = total
        # the dominant kernel (largest share of the step's time): mean DRAM bytes per launch, what roofline.traffic reports
        by_name = {}
        for rec in out_rows:
            b = by_name.setdefault(rec['Kernel Name'], [0.0, 0.0, 0])
            b[0] += rec.get('gpu__time_duration.sum', 0)
            b[1] += rec.get('dram__bytes_read.sum', 0) + rec.get('dram__bytes_write.sum', 0)
            b[2] += 1
        dom = max(by_name.items(), key=lambda kv: kv[1][0])
        traffic[workload][key] = dom[1][1] / dom[1][2]
        print(f'   dominant kernel {dom[0][:60]}: {dom[1][2]} launches, {dom[1][1] / dom[1][2]:.4g} B per launch')
    traffic['_source'] = ('DRAM bytes of ONE STEP (all kernels of the step): dram__bytes_read.sum + dram__bytes_write.sum from '
                          '`ncu --set full` captures, profiles/<tag>_ncu_full_<name>.csv (profiles/tools/ncu_round2.sh) under '
                          '"per_step"; the plain entries are per LAUNCH of the step\'s dominant kernel (one RK stage for heat2d)')
    json.dump(traffic, open(traffic_path, 'w'), indent=1)


if __name__ == '__main__':
    main()
