#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + top stall sites (SASS).  Usage:
    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [n_top]"""
import csv
import subprocess
import sys
import io

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    h, units = rows[0], rows[1]
    print('kernels:', [r[h.index('Kernel Name')][:60] for r in rows[2:]])
    for w in WANT:
        for i, x in enumerate(h):
            if x == w:
                print('%-70s %-12s %s' % (w, units[i], [r[i] for r in rows[2:]]))
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv', '--print-source', 'sass']))))
    blocks = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    start = blocks[0]
    end = blocks[1] if len(blocks) > 1 else len(rows)
    h = rows[start + 1]
    col = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    tot = dict.fromkeys(stalls, 0)
    samples, agg = 0, []
    for r in rows[start + 2:end]:
        if len(r) < len(h):
            continue
        ns = int(r[col['# Samples']] or 0)
        samples += ns
        d = {}
        for s in stalls:
            v = int(r[col[s]] or 0)
            tot[s] += v
            if v:
                d[s[6:]] = v
        agg.append((ns, r[col['Address']][-5:], r[col['Source']][:70], d))
    print('total samples', samples)
    for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
        print('  %-28s %8d %5.1f%%' % (s, v, 100.0 * v / max(samples, 1)))
    agg.sort(key=lambda t: -t[0])
    for a in agg[:ntop]:
        print(a[0], a[1], a[2], a[3])


if __name__ == '__main__':
    main()
