#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into the handful of numbers the roofline
discussion needs, plus the top stall sites of the source page.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/ncu_<kernel>_rNN.md
    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep --traffic-key conv_chain_bd4 [--launch K]
        additionally records dram read+write bytes of launch K (default: the longest launch) under that
        key in profiles/ncu_traffic.json -- the file bench.py reads `roofline.traffic` from.
"""
import json
import os
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
]


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def _bytes(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def record_traffic(rep, hdr, units, rows, key, launch):
    ird, iwr, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
    if launch is None:
        launch = max(range(len(rows)), key=lambda k: float(rows[k][it].replace(',', '')))
    r = rows[launch]
    ent = {'dram_bytes_per_launch': _bytes(r[ird], units[ird]) + _bytes(r[iwr], units[iwr]),
           'dram_read_bytes': _bytes(r[ird], units[ird]), 'dram_write_bytes': _bytes(r[iwr], units[iwr]),
           'kernel': r[hdr.index('Kernel Name')][:80], 'launch': launch,
           'src': f'ncu --set full capture {os.path.basename(rep)} (see profiles/)'}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')
    data = json.load(open(path)) if os.path.isfile(path) else {}
    data[key] = ent
    json.dump(data, open(path, 'w'), indent=1, sort_keys=True)


def main(rep, traffic_key=None, launch=None):
    raw = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    if traffic_key:
        record_traffic(rep, hdr, units, rows, traffic_key, launch)
    print(f'# ncu --set full summary: {rep}\n')
    name_i = hdr.index('Kernel Name')
    for k, r in enumerate(rows):
        print(f'## launch {k}: `{r[name_i][:90]}`\n')
        print('| metric | value | unit |\n|---|---:|---|')
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f'| {key} | {r[i]} | {units[i]} |')
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')].replace(',', ''))
            wr = float(r[hdr.index('dram__bytes_write.sum')].replace(',', ''))
            print(f'\ntraffic = dram read {rd} {units[hdr.index("dram__bytes_read.sum")]} + write {wr} '
                  f'{units[hdr.index("dram__bytes_write.sum")]}\n')
        except Exception:
            print()
    src = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv']))))
    h = None
    for i, r in enumerate(src):
        if 'Source' in r and '# Samples' in r:
            h = i
            break
    if h is None:
        return
    hdr = src[h]
    isrc, isamp = hdr.index('Source'), hdr.index('# Samples')
    stall = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
    data = []
    for r in src[h + 1:]:
        try:
            data.append((int(r[isamp]), r))
        except Exception:
            pass
    tot = sum(s for s, _ in data) or 1
    agg = {}
    for s, r in data:
        for i in stall:
            if r[i] not in ('', '0'):
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print('## warp-state samples (first launch)\n')
    print(', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print('\n| samples | share | SASS | top stall |\n|---:|---:|---|---|')
    for s, r in sorted(data, key=lambda x: -x[0])[:14]:
        st = sorted(((hdr[i], int(r[i])) for i in stall if r[i] not in ('', '0')), key=lambda kv: -kv[1])[:2]
        print(f'| {s} | {100 * s / tot:.1f}% | `{r[isrc].strip()[:60]}` | {st} |')


if __name__ == '__main__':
    a = sys.argv[1:]
    key = a[a.index('--traffic-key') + 1] if '--traffic-key' in a else None
    lk = int(a[a.index('--launch') + 1]) if '--launch' in a else None
    main(a[0], key, lk)
