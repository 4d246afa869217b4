#!/usr/bin/env python
"""Blackwell-specific SASS mnemonics per kernel of the built library (evidence that the hot kernels are tcgen05 / TMEM / TMA code):

    python tools/sass_grep.py > profiles/round2_sass_grep.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mb_istft_vits_b200", "libmbistft.so")
PAT = re.compile(r"\b(UTCHMMA[.\w]*|UTCQMMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAPF[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|"
                 r"UCGABAR_\w+|SYNCS[.\w]*|MUFU\.TANH|MUFU\.RCP|LDG\.E\.ENL2\.256|STG\.E\.ENL2\.256|FFMA2|FADD2|FMUL2)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur:
            for tok in PAT.findall(line):
                per[cur][tok] += 1
    names = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("# cuobjdump -sass mb_istft_vits_b200/libmbistft.so, Blackwell-specific mnemonics per kernel (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)")
    print("# tcgen05.mma -> UTCHMMA (.2CTA for cta_group::2); tcgen05.ld -> LDTM; TMA -> UTMALDG (UTMAPF = L2 prefetch); tcgen05.commit -> UTCBAR;")
    print("# mbarrier -> SYNCS; 256-bit global accesses -> LDG/STG.E.ENL2.256; packed fp32 pairs -> FFMA2 / FADD2 / FMUL2")
    print("# totals over the library: " + ", ".join("%s %d" % kv for kv in sorted(total.items())))
    print()
    for (k, c), n in zip(per.items(), names):
        if not c:
            continue
        n = re.sub(r"\(.*", "", n)
        print(n)
        print("    " + ", ".join("%s %d" % kv for kv in sorted(c.items())))


if __name__ == "__main__":
    main()
