#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show which Blackwell / Hopper-class hardware paths the built library
uses (cuobjdump -sass of stochy.jl_b200/libstochy_b200.so; runs without a GPU):

  UBLKCP    TMA bulk copy (cp.async.bulk.shared::cluster.global)        SYNCS   mbarrier (arrive / try_wait / expect_tx)
  LDGSTS    cp.async (4/8-byte tile records)                            UCGABAR thread-block-cluster barrier
  REDUX     warp reduce in one instruction                              CCTL / ERRBAR / MEMBAR   fences
  ACQBULK / UTMAPF  bulk-copy proxy ops / L2 prefetch                    BAR     CTA barriers

  python scripts/sass_excerpt.py > profiles/r02_sass_excerpt.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stochy.jl_b200", "libstochy_b200.so")
WATCH = ["UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "UCGABAR", "CGAERRBAR", "ACQBULK", "REDUX", "BAR", "MEMBAR", "ERRBAR", "ATOMG", "ATOMS", "RED",
         "SHFL", "VOTE", "IMAD", "DFMA", "MUFU", "HMMA", "UTCMMA", "STG", "LDG", "LDS", "STS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts, total, arch = None, {}, {}, None
    for line in out.splitlines():
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            counts[kern] = collections.Counter(); total[kern] = 0
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            op = m.group(1)
            total[kern] += 1
            for w in WATCH:                                      # UCGABAR_ARV / UCGABAR_WAIT count as UCGABAR
                if op == w or op.startswith(w + "_"):
                    counts[kern][w] += 1
    print("# SASS mnemonic counts per kernel of libstochy_b200.so (%s; static instruction counts)" % arch)
    print("# %-78s %6s  %s" % ("kernel", "instr", "  ".join(WATCH[:10])))
    for k in sorted(counts, key=lambda k: -total[k]):
        short = re.sub(r"\((int|bool)\)", "", k.rsplit(">(", 1)[0] + (">" if ">(" in k else "")).replace("void ", "")[:78]
        c = counts[k]
        sel = "  ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w])
        print("%-80s %6d  %s" % (short, total[k], sel))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("\n# whole library: " + "  ".join("%s=%d" % (w, tot[w]) for w in WATCH if tot[w]))
    print("# tensor-core instructions (HMMA / UTCMMA / tcgen05): %d -- nothing on this path is a dense contraction" % (tot["HMMA"] + tot["UTCMMA"]))


if __name__ == "__main__":
    main()
