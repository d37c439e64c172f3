#!/usr/bin/env python
"""Summaries of the ncu captures of tools/profile_r02.sh for profiles/:
    python tools/ncu_summarize.py launches gpurun_out/prof2/launches.csv.gz  > profiles/r02_ncu_launch_shares.txt
    python tools/ncu_summarize.py full     gpurun_out/prof2/full_a_raw.csv   > profiles/r02_ncu_full_summary.md
"""
import csv
import gzip
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
    name = re.sub(r'^void ', '', name)
    m = re.match(r'([A-Za-z0-9_:]+(<[^()]*>)?)', name)
    return (m.group(1) if m else name)[:90]


def launches(path):
    rows = list(csv.reader(gzip.open(path, 'rt') if path.endswith('.gz') else open(path)))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hdr]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg, order = {}, []
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        k = short(r[ki])
        ms = float(r[vi].replace(',', '')) * scale.get(r[ui], 1e-6)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if not k.startswith('at::'))
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms of kernel time "
          f"({ours:.2f} ms in this library's kernels); per-launch times are serialised and cold-cache: compare SHARES")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  {k}")


METRICS = OrderedDict([
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
    ('FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
    ('lts__t_sector_hit_rate.pct', 'L2hit%'),
    ('l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed', 'smem_rd%'),
    ('l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed', 'smem_wr%'),
])


def full(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    stall_cols = [(n, i) for n, i in col.items() if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio')]
    print("| kernel | grid | block | time µs | DRAM GB (r+w) | DRAM TB/s | DRAM % | tensor pipe % | SM % | warps active % | L2 hit % | smem rd/wr % | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    def num(r, name):
        i = col.get(name)
        if i is None or i >= len(r):
            return None
        try:
            return float(r[i].replace(',', ''))
        except ValueError:
            return None
    def to_s(r, name):
        v = num(r, name)
        u = units[col[name]] if name in col else ''
        if v is None:
            return None
        return v * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 's': 1.0, 'second': 1.0}.get(u, 1e-9)
    def to_b(r, name):
        v = num(r, name)
        u = units[col[name]] if name in col else ''
        if v is None:
            return 0.0
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)
    for r in rows[2:]:
        if len(r) < 10:
            continue
        name = short(r[col['Kernel Name']])
        t = to_s(r, 'gpu__time_duration.sum')
        byts = to_b(r, 'dram__bytes_read.sum') + to_b(r, 'dram__bytes_write.sum')
        stalls = sorted(((num(r, n) or 0.0, n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for n, _ in stall_cols), reverse=True)
        tot = sum(s for s, _ in stalls) or 1.0
        top = ', '.join(f"{n} {100 * s / tot:.0f}%" for s, n in stalls[:3])
        g = lambda k: (f"{num(r, k):.1f}" if num(r, k) is not None else '-')
        print(f"| `{name}` | {r[col['Grid Size']]} | {r[col['Block Size']]} | {t * 1e6:.1f} | {byts / 1e9:.3f} | {byts / t / 1e12:.2f} | "
              f"{g('FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed')} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} | "
              f"{g('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | {g('sm__warps_active.avg.pct_of_peak_sustained_active')} | {g('lts__t_sector_hit_rate.pct')} | "
              f"{g('l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed')}/{g('l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed')} | {top} |")


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
