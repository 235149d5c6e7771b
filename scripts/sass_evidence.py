#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md "What proves a
Blackwell-native kernel"): UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UBLKCP (TMA / bulk copies), REDG
(vector reductions), LDGSTS (cp.async), LDGMC / STG...MC / REDG.MC (multimem.* through the NVSwitch).

    python scripts/sass_evidence.py > profiles/r2_sass_evidence.txt      (runs cuobjdump -sass on the built library)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hhfm_b200", "libhhfm_sm100.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|REDG[.\w]*|LDGSTS[.\w]*|"
                 r"UTCBAR[.\w]*|LDGMC[.\w]*|STG[.\w]*\.MC[.\w]*|[A-Z]+\.MC\.[.\w]*)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        if cur:
            for t in PAT.findall(ln):
                cnt[cur][t] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass hhfm_b200/libhhfm_sm100.so (sm_100a), mnemonic counts per kernel; kernels without any are omitted")
    for (k, c), name in zip(cnt.items(), names):
        if not c:
            continue
        name = re.sub(r"\(.*", "", name)
        print(name)
        print("    " + ", ".join("%s x%d" % (t, n) for t, n in sorted(c.items())))


if __name__ == "__main__":
    sys.exit(main())
