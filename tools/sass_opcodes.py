#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor copies), UBLKCP (bulk
copies), HMMA (mma.sync), LDSM / STSM (ldmatrix / stmatrix), SYNCS (mbarrier), FFMA2 / FADD2 (packed fp32).

    python tools/sass_opcodes.py [path/to/lib.so] > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "crop2seg_b200", "lib", "libcrop2seg_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDSM", "STSM", "SYNCS", "FFMA2",
         "FADD2", "MUFU", "LDG", "STG", "LDS", "STS", "RED", "ATOM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur, arch = collections.OrderedDict(), None, set()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    demangled = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: arch {sorted(arch)}; SASS instruction counts per kernel (static, not executed)")
    print("# columns: total | " + " ".join(WATCH))
    for (name, cnt), nice in zip(kernels.items(), demangled):
        nice = nice.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")
        nice = re.sub(r"^void ", "", nice)
        nice = re.sub(r"\((?:const |CUtensorMap|c2s::|float|int|unsigned|long|void|__nv|uint|TileArgs|ClassArgs).*$", "", nice)[:110]
        hot = " ".join(f"{w}={cnt[w]}" for w in WATCH if cnt[w])
        print(f"{nice:110s} total={cnt['total']:6d} {hot}")


if __name__ == "__main__":
    main()
