"""Count the SASS mnemonics that prove the Blackwell data path per kernel of lvllm_b200/libb200moe.so (cuobjdump -sass):
tcgen05 MMAs (UTCHMMA / UTCQMMA / UTCOMMA), tensor-map TMA (UTMALDG), bulk async copies (UBLKCP), cp.async (LDGSTS), TMEM
loads / stores (LDTM / STTM), and the legacy HMMA that must NOT appear.  usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lvllm_b200", "libb200moe.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|UTCIMMA|UTMALDG|UTMASTG|UBLKCP|LDGSTS|LDTM|STTM|UTCBAR|UTCCP|HMMA|IMMA|QMMA|F2FP)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = PAT.search(ln)
        if m:
            per[cur][m.group(1)] += 1
    try:
        names = subprocess.run(["c++filt"], input="\n".join(per), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    except OSError:
        names = list(per)
    tot = collections.Counter()
    print(f"# SASS evidence: cuobjdump -sass {os.path.relpath(LIB, ROOT)} ({len(per)} kernels); counts of instruction SITES per kernel")
    for (mangled, c), nm in zip(per.items(), names):
        tot.update(c)
        short = re.sub(r"\(.*", "", nm).replace("void b200::", "")
        print(f"{short:90s} {dict(c)}")
    print("# totals:", dict(tot))
    print("# legacy mma.sync sites (HMMA / IMMA / QMMA), must be 0:", tot["HMMA"] + tot["IMMA"] + tot["QMMA"])


if __name__ == "__main__":
    sys.exit(main())
