#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tensor-core / TMA / TMEM / bulk-copy paths, from the
in-tree libcfm_b200.so (no GPU needed: cuobjdump disassembles the embedded sm_100a cubins).

    python scripts/sass_evidence.py > profiles/r02_sass_evidence.txt

UTCHMMA / UTCQMMA ... = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld (TMEM read-out),
UTCBAR = tcgen05.commit -> mbarrier, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, REDUX = warp reductions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cfm_b200", "libcfm_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP",
       "SYNCS", "REDUX", "MUFU.EX2", "HMMA", "FFMA", "DFMA", "BAR.SYNC", "ATOMG", "REDG", "LDGSTS", "CCTL"]


def main():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in PAT:
            if re.search(r"(?<![A-Z])" + re.escape(p) + r"\b", ln):
                counts[cur][p] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS evidence from {os.path.relpath(LIB, ROOT)} (cuobjdump -sass), sm_100a")
    print("# kernel | " + " ".join(PAT))
    for (name, c), dn in zip(counts.items(), demangle):
        if not any(c.values()):
            continue
        short = re.sub(r"\(.*", "", dn)[:110]
        print(f"{short:<112} | " + " ".join(f"{p}={c[p]}" for p in PAT if c[p]))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("# totals: " + " ".join(f"{p}={tot[p]}" for p in PAT if tot[p]))


if __name__ == "__main__":
    sys.exit(main())
