"""Summarise an ncu report exported with `--page raw --csv` (+ optionally `--page source --csv --print-source sass`)."""
import collections
import csv
import re
import sys

WANT = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_branch_resolving', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_no_instructions',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_sample_buffer_overflow']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernels:", [r[idx['Kernel Name']][:40] for r in data])
    for w in WANT:
        if w in idx:
            print(f"{w:80s} {units[idx[w]]:10s}", [r[idx[w]][:18] for r in data])


def source(path, top=24):
    rows = list(csv.reader(open(path)))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            kern.append(cur)
            continue
        if cur is None:
            continue
        if cur['hdr'] is None:
            cur['hdr'] = r
            continue
        cur['rows'].append(r)
    for k in kern:
        h = {n: i for i, n in enumerate(k['hdr'])}
        tot = thr = 0
        byop = collections.Counter()
        for r in k['rows']:
            try:
                n, t = int(r[h['Instructions Executed']]), int(r[h['Thread Instructions Executed']])
            except Exception:
                continue
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[h['Source']])
            byop[m.group(2).split('.')[0] if m else '?'] += n
            tot += n
            thr += t
        print(k['name'][:70], 'warp-inst', tot, 'avg active threads', round(thr / max(tot, 1), 2))
        print('   ' + '  '.join(f'{op} {n / tot * 100:.1f}%' for op, n in byop.most_common(top)))


if __name__ == '__main__':
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[2])
