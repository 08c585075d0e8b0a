#!/usr/bin/env python
"""Per-kernel SASS extract of libtreasure_b200.so (cuobjdump): code size, and the counts of the mnemonics that show
what the kernels are built from -- UBLKCP (cp.async.bulk, the TMA engine's 1-D bulk copies), SYNCS (mbarrier),
REDUX (warp reductions), ATOMS / ATOMG / RED (atomics), BAR, CALL (out-of-line device functions), LDL / STL (local
memory: spills), IMAD.WIDE / IMAD.HI (Philox).  usage: sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gym_treasure_game_b200", "libtreasure_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, stats = None, collections.OrderedDict()
KEYS = ["UBLKCP", "SYNCS", "REDUX", "ATOMS", "ATOMG", "RED", "BAR", "CALL", "LDL", "STL", "IMAD.WIDE", "IMAD.HI", "VOTE", "SHFL", "LDS", "STS", "LDG", "STG"]
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); stats[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        stats[cur]["_n"] += 1
        op = m.group(2)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                stats[cur][k] += 1
demangle = subprocess.run(["c++filt"] + list(stats), capture_output=True, text=True).stdout.splitlines()
print("%-8s %-9s %s" % ("instr", "KB", "kernel  [mnemonic counts]"))
for (name, c), dn in zip(stats.items(), demangle):
    short = re.sub(r"\(.*", "", dn)
    print("%-8d %-9.1f %s  [%s]" % (c["_n"], c["_n"] * 16 / 1024, short, ", ".join("%s %d" % (k, c[k]) for k in KEYS if c[k])))
