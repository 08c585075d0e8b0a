#!/usr/bin/env python
"""Aggregates `ncu -i rep --page source --csv --print-source cuda,sass` by source region: the regions are the
device functions of csrc/*.cu(h) (definition line up to the next definition) and, inside kernels, the
`// ---- phase` markers.  usage: ncu_regions.py file.csv [min_pct]"""
import csv, collections, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gym_treasure_game_b200", "csrc")


def num(x):
    try: return int(x)
    except Exception: return 0


def regions_of(path):
    out = []
    fn = re.compile(r'^(?:static\s+)?(?:__device__|__global__|tg_\w+_kernel)\b.*?(\w+)\s*\(')
    pend = None
    for ln, line in enumerate(open(path), 1):
        if line.startswith("template"): pend = ln; continue
        m = fn.match(line)
        if m:
            name = m.group(1)
            if name in ("__launch_bounds__",):
                continue
            out.append((pend or ln, name)); pend = None; continue
        m = re.match(r'^tg_(\w+)\(', line)
        if m: out.append((pend or ln, "tg_" + m.group(1))); pend = None; continue
        m = re.search(r'// ---- (phase \w+|sort)', line)
        if m and out: out.append((ln, out[-1][1].split(':')[0] + ": " + m.group(1)))
        if not line.startswith("template") and line.strip() and not line.startswith("__"): pend = None if not line.startswith("template") else pend
    return out


def main(path, min_pct=0.4):
    regs = {f: regions_of(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))}
    rd = csv.reader(open(path))
    cur = None
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
    for row in rd:
        if not row: continue
        if row[0] == "File Path": cur = row[1].split('/')[-1]; continue
        if row[0] == "Function Name": continue
        if row[0] == "Line No":
            h = row; iI = h.index("Instructions Executed"); iT = h.index("Thread Instructions Executed"); iS = h.index("# Samples")
            iN = h.index("stall_no_inst"); iL = h.index("stall_long_sb"); iB = h.index("stall_barrier"); continue
        if row[0] == "": continue
        try: ln = int(row[0])
        except Exception: continue
        name = cur
        for (start, nm) in regs.get(cur, []):
            if start <= ln: name = nm
            else: break
        a = agg[name]
        a[0] += num(row[iI]); a[1] += num(row[iT]); a[2] += num(row[iS]); a[3] += num(row[iN]); a[4] += num(row[iL]); a[5] += num(row[iB])
    tot = [sum(v[k] for v in agg.values()) for k in range(6)]
    print("total inst %d thread-inst %d samples %d noinst %d longsb %d barrier %d" % tuple(tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if 100.0 * v[0] / max(tot[0], 1) < min_pct and 100.0 * v[2] / max(tot[2], 1) < min_pct: continue
        print("%-34s inst %9d (%4.1f%%) lanes %5.1f samples %6d (%4.1f%%) noinst %6d longsb %6d barrier %6d" % (
            k, v[0], 100 * v[0] / max(tot[0], 1), v[1] / max(v[0], 1), v[2], 100 * v[2] / max(tot[2], 1), v[3], v[4], v[5]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.4)
