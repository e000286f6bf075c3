"""Where the warp instructions of the traversal kernels go: node loop vs everything else, from an ncu source page exported
with `ncu -i rep --page source --csv --print-source sass`.  python tools/ncu_sections.py src.csv"""
import csv
import sys


def kernels(path):
    rows = list(csv.reader(open(path)))
    out, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'hdr': None, 'rows': []}
            out.append(cur)
            continue
        if cur is None:
            continue
        if cur['hdr'] is None:
            cur['hdr'] = r
            continue
        cur['rows'].append(r)
    seen, uniq = set(), []
    for k in out:          # the export lists every kernel twice (source and SASS view)
        if k['name'] not in seen:
            seen.add(k['name'])
            uniq.append(k)
    return uniq


def main(path):
    for k in kernels(path):
        if 'k_trace' not in k['name']:
            continue
        h = {n: i for i, n in enumerate(k['hdr'])}
        def num(r, c):
            try:
                return int(r[h[c]])
            except Exception:
                return 0
        R = k['rows']
        n = [num(r, 'Instructions Executed') for r in R]
        t = [num(r, 'Thread Instructions Executed') for r in R]
        s = [num(r, 'Warp Stall Sampling (All Samples)') for r in R]
        src = [r[h['Source']] for r in R]
        tot, tots = sum(n), sum(s)
        i0 = next(i for i, x in enumerate(src) if 'FFMA2' in x)
        a = i0
        while a > 0 and 'LDC.64' not in src[a]:
            a -= 1
        b = i0
        while b < len(R) - 1 and not ('BRA' in src[b] and n[b] == n[i0] and 'ISETP.GT.AND' in src[b - 1]):
            b += 1
        loop_n, loop_t, loop_s = sum(n[a:b + 1]), sum(t[a:b + 1]), sum(s[a:b + 1])
        visits = n[i0]
        print(k['name'][:60])
        print(f"  warp instructions {tot:,}; node loop (SASS lines {a}..{b}): {loop_n / tot * 100:.1f} % of them, {loop_t / max(loop_n, 1):.1f} of 32 lanes active, "
              f"{loop_s / tots * 100:.1f} % of the stall samples; {visits:,} warp-level node visits = {loop_n / visits:.1f} instructions per visit")
        print(f"  rest (ray set-up, instance entry, element tests, hit output): {(tot - loop_n) / tot * 100:.1f} % of the instructions, "
              f"{(sum(t) - loop_t) / max(tot - loop_n, 1):.1f} lanes active, {(tots - loop_s) / tots * 100:.1f} % of the samples")
        first_use = s[i0]
        print(f"  first use of the node record (line {i0}): {first_use / tots * 100:.1f} % of all stall samples = exposed load latency")


if __name__ == '__main__':
    main(sys.argv[1])
