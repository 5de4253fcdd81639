#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass): DMMA (FP64 tensor core), UTMALDG (TMA loads),
SYNCS (mbarrier), SHFL, LDS/STS, DFMA ... -> profiles/rNN_sass_counts.txt.  Evidence tooling, not product code."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vbmatrixfactorization.jl_b200", "lib", "libvbmf_b200.so")
OPS = ["DMMA", "UTMALDG", "SYNCS", "SHFL", "LDS", "STS", "LDG", "STG", "DFMA", "DMUL", "DADD", "MUFU", "BAR", "ATOM", "RED"]


def main(out):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    counts, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            counts[name][m.group(1).split(".")[0]] += 1
    rows = []
    for n, c in counts.items():
        d = demangle(n)
        d = re.sub(r"\(.*", "", d).replace("void vb::", "").replace("vb::", "")
        rows.append((d, sum(c.values()), [c.get(o, 0) for o in OPS]))
    rows.sort(key=lambda r: -r[2][0])
    with open(out, "w") as f:
        f.write("# SASS mnemonic counts per kernel of vbmatrixfactorization.jl_b200/lib/libvbmf_b200.so (sm_100a), `cuobjdump -sass`\n")
        f.write("# DMMA = mma.sync.m8n8k4.f64 (the FP64 tensor-core instruction; tcgen05 has no f64 kind), UTMALDG = cp.async.bulk.tensor (TMA),\n")
        f.write("# SYNCS = mbarrier arrive/try_wait.  No ATOM/RED anywhere on the path (fixed-order reductions).\n")
        f.write("%-64s %7s " % ("kernel", "instrs") + " ".join("%7s" % o for o in OPS) + "\n")
        for d, tot, v in rows:
            f.write("%-64s %7d " % (d[:64], tot) + " ".join("%7d" % x for x in v) + "\n")
        tot = [sum(r[2][i] for r in rows) for i in range(len(OPS))]
        f.write("%-64s %7d " % ("TOTAL (%d kernels)" % len(rows), sum(r[1] for r in rows)) + " ".join("%7d" % x for x in tot) + "\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_counts.txt"))
