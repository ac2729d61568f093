#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep profiles/X_full.md [profiles/traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct']


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if not l.startswith('==')))
    hdr = rows[0]
    iK, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iV:
            continue
        v = float(r[iV].replace(',', ''))
        v = v / 1e3 if r[iU] == 'ns' else v * 1e3 if r[iU] == 'ms' else v
        agg.setdefault(r[iK], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, 'w') as f:
        f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n')
        f.write('source: %s, %d launches, %.1f us total\n\n| kernel | launches | mean us | share |\n|---|---|---|---|\n'
                % (src, sum(len(v) for v in agg.values()), tot))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write('| `%s` | %d | %.2f | %.1f%% |\n' % (k[:110], len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print(open(dst).read())


def full(src, dst, traffic=None):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ['# ncu --set full summary of %s\n' % src]
    per_launch = []
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        out.append('\n## %s\n\n| metric | value | unit |\n|---|---|---|' % name)
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append('| %s | %s | %s |' % (k, r[i], units[i]))
                vals[k] = (r[i], units[i])
        per_launch.append(vals)
    # hot regions of the first profiled launch from the source page
    srcp = subprocess.run(['ncu', '-i', src, '--page', 'source', '--csv', '--print-source', 'sass'],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(srcp.splitlines()))
    if len(rows) > 3:
        hdr = rows[1]
        iA, iS, iI, iN = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
        d, seen = [], set()
        for r in rows[2:]:
            if len(r) <= iI or not r[iI].isdigit():
                continue
            a = int(r[iA], 16)
            if a in seen:
                break
            seen.add(a)
            d.append((a, r[iS].strip(), int(r[iI]), int(r[iN])))
        tot = sum(x[2] for x in d)
        tots = sum(x[3] for x in d)
        out.append('\n## SASS regions of the first launch (consecutive instructions with the same execution count)\n')
        out.append('total warp instructions %d, %d SASS instructions, %d stall samples\n' % (tot, len(d), tots))
        out.append('| SASS range | instrs | exec per instr | warp-instr | share | stall samples | first instruction |\n|---|---|---|---|---|---|---|')
        cur, start, acc, accs = None, 0, 0, 0
        segs = []
        for k, (a, s, i, n) in enumerate(d):
            if cur is None:
                cur, start = i, k
            if abs(i - cur) > 0.02 * max(cur, 1) + 8:
                segs.append((start, k - 1, cur, acc, accs))
                cur, start, acc, accs = i, k, 0, 0
            acc += i
            accs += n
        segs.append((start, len(d) - 1, cur, acc, accs))
        for s0, s1, c, acc, accs in segs:
            if acc > 0.004 * tot or accs > 0.01 * tots:
                out.append('| %#x-%#x | %d | %d | %d | %.1f%% | %d | `%s` |' % (
                    d[s0][0] - d[0][0], d[s1][0] - d[0][0], s1 - s0 + 1, c, acc, 100.0 * acc / tot, accs, d[s0][1][:48]))
    open(dst, 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))
    if traffic and per_launch:
        def num(v):
            x, u = v
            x = float(x.replace(',', ''))
            return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        t = [num(v['dram__bytes_read.sum']) + num(v['dram__bytes_write.sum']) for v in per_launch]
        json.dump({'dram_bytes_per_launch': sum(t) / len(t), 'launches': len(t), 'source': src,
                   'note': 'dram__bytes_read.sum + dram__bytes_write.sum of k_model_step, ncu --set full'},
                  open(traffic, 'w'))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
