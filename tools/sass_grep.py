"""Counts of the Blackwell-native SASS mnemonics per kernel of libpio_sm100.so (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor
load/store/reduce, UBLKCP = bulk copy, HMMA = legacy mma.sync.

    python tools/sass_grep.py > profiles/r02_sass_grep.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "patch-ioner_b200", "libpio_sm100.so")
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UTCBAR|UTCCP|HMMA|SYNCS|ATOM|RED|LDGSTS)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        per[cur]["_instructions"] += 1 if re.search(r"/\*[0-9a-f]{4,}\*/", line) else 0
        m = PAT.search(line)
        if m:
            name = m.group(1)
            if name == "UTCHMMA" and ".2CTA" in line:
                name = "UTCHMMA.2CTA"
            per[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({len(per)} kernels)")
    tot = collections.Counter()
    for (k, c), d in zip(per.items(), demangle):
        short = re.sub(r"^void ", "", d.replace("(anonymous namespace)::", ""))
        short = re.sub(r"\((?:CUtensorMap_st|float|int|void|__nv|unsigned|pio::|long|char|const)[^)]*.*$", "", short)[:70]
        keys = [n for n in c if n != "_instructions" and n not in ("SYNCS", "ATOM", "RED", "LDGSTS")]
        if not keys:
            continue
        tot.update({n: c[n] for n in keys})
        print(f"{short:70s} instr {c['_instructions']:6d}  " + "  ".join(f"{n}={c[n]}" for n in sorted(keys)))
    print("# totals: " + "  ".join(f"{n}={v}" for n, v in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
