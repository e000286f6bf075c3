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


JSON_KEYS = {
    "duration_ms": ("gpu__time_duration.sum", 1.0), "warp_instructions": ("smsp__inst_executed.sum", 1.0),
    "lanes_per_instruction": ("smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "l1_lsu_wavefronts_pct": ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0),
    "alu_pipe_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
    "fma_pipe_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
    "l1_hit_pct": ("l1tex__t_sector_hit_rate.pct", 1.0), "registers": ("launch__registers_per_thread", 1.0),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    "dram_read_bytes": ("dram__bytes_read.sum", None), "dram_write_bytes": ("dram__bytes_write.sum", None),
    "l1_bytes": ("l1tex__t_bytes.sum", None), "l2_bytes": ("lts__t_bytes.sum", None),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}


def to_json(path, out):
    """per-kernel numbers of a `--page raw --csv` export as JSON (what bench.py's roofline record cites)."""
    import json
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    res = {"source": "ncu --set full --clock-control none, whole-frame launches of the headline config (1920x1080, 16 spp), second frame"}
    for r in data:
        name = re.sub(r"^void\s+", "", r[idx["Kernel Name"]])
        name = re.sub(r"^yrt::", "", name).split("(")[0].split("<")[0]
        d = {}
        for k, (m, _) in JSON_KEYS.items():
            if m in idx and r[idx[m]] not in ("", "n/a"):
                v = float(r[idx[m]].replace(",", ""))
                u = units[idx[m]]
                if u in UNIT and (k.endswith("bytes") or k == "duration_ms"):
                    v *= UNIT[u]
                d[k] = v
        if "dram_read_bytes" in d and "dram_write_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
        res[name] = d
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


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
    if sys.argv[1] == '--json':          # python tools/ncu_summary.py --json raw.csv out.json
        to_json(sys.argv[2], sys.argv[3])
    else:
        raw(sys.argv[1])
        if len(sys.argv) > 2:
            source(sys.argv[2])
