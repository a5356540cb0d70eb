#!/usr/bin/env python
"""Counts the Blackwell-specific SASS opcodes per kernel in the built libdeepcut_b200.so (cuobjdump -sass): the evidence that
the convolution kernel is tcgen05 / TMEM / TMA code (B200_PROFILING.md lists the mnemonics).  Writes a markdown table.

  python tools/sass_counts.py > profiles/r2_sass_counts.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deepcut-cnn_b200", "libdeepcut_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.*2CTA", "UTMASTG", "LDTM", "UTCBAR", "UTCBAR.*MULTICAST", "UTCATOMSWS", "LDGSTS", "FHFMA", "SYNCS", "R2UR.BROADCAST", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        for op in OPS:
            if re.search(r"\b" + op.replace(".", r"\.").replace(r"\.*", ".*") + r"\b", line):
                cur[op] += 1
    print("# SASS opcode counts, libdeepcut_b200.so (`cuobjdump -sass`, sm_100a)\n")
    print("| kernel | " + " | ".join(OPS) + " |")
    print("|---|" + "---|" * len(OPS))
    tot = collections.Counter()
    for name, c in per.items():
        if not any(c.values()):
            continue
        d = demangle(name)
        d = re.sub(r"\(CUtensorMap.*", "", d).replace("void dc::", "").replace("dc::", "").replace("(int)", "")
        print("| `%s` | " % d + " | ".join(str(c[o]) for o in OPS) + " |")
        tot.update(c)
    print("| **all kernels** | " + " | ".join(str(tot[o]) for o in OPS) + " |")
    print("\nUTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA bulk-tensor load / store, LDTM = tcgen05.ld "
          "(TMEM -> registers), UTCBAR = tcgen05.commit (mbarrier arrive, MULTICAST across the CTA pair), LDGSTS = cp.async, "
          "FHFMA = mixed fp16 x fp16 + fp32 FMA of the epilogue, SYNCS = mbarrier try_wait / arrive.  R2UR.BROADCAST / ELECT: ptxas's "
          "per-lane waterfall around a uniform-datapath instruction whose operands it cannot prove warp-uniform (round 1 issued TMA loads "
          "and MMAs under `lane == 0`: 599 R2UR.BROADCAST in the library; with an elect.sync lane of a converged warp none are left in the "
          "producer / MMA warps -- the remaining ELECTs are the elect.sync themselves and the epilogue's TMA stores; profiles/r2_issue_lane.md).")


if __name__ == "__main__":
    main()
