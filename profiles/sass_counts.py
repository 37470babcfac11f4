"""Per-kernel SASS instruction counts of the built library (structure item 15): runs `cuobjdump -sass` on
atlas_b200/libsptrans_b200.so and counts, for every kernel, the mnemonics that prove which hardware path it uses
(B200_PROFILING.md): DMMA (fp64 tensor pipe), UTCHMMA / UTCQMMA / UTCIMMA (tcgen05.mma), LDTM / STTM (tensor memory),
UBLKCP (cp.async.bulk, TMA engine, 1-D), UTMALDG / UTMASTG (tensor-map TMA), LDGSTS (cp.async), SYNCS (mbarrier),
DFMA / DADD / DMUL (fp64 pipe), LDS / STS, STL / LDL (spills).  Usage: python profiles/sass_counts.py > profiles/sass_counts_r02.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "atlas_b200", "libsptrans_b200.so")
KEYS = ["DMMA", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS",
        "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "STL", "LDL", "BAR", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", out)
    if names:
        d = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout.splitlines()
        demangle = dict(zip(names, d))
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            cur = cur.replace("(anonymous namespace)::", "")
            cur = re.sub(r"\(.*", "", cur)   # drop the argument list
            while cur in counts:              # (overloads / identical names from different translation units)
                cur += "'"
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, REPO)} (sm_100a): instructions per kernel, static counts")
    print("# kernel | total | " + " | ".join(KEYS))
    tot = collections.Counter()
    for k, c in counts.items():
        tot.update(c)
        print(k + " | " + str(c["total"]) + " | " + " | ".join(str(c[x]) for x in KEYS))
    print("ALL KERNELS | " + str(tot["total"]) + " | " + " | ".join(str(tot[x]) for x in KEYS))
    syms = subprocess.run(["nm", "-D", "--undefined-only", LIB], stdout=subprocess.PIPE, text=True).stdout
    libs = sorted({s for s in ("cublas", "cufft", "nccl", "cudnn", "cusolver") if s in syms.lower()})
    print("# undefined symbols from cuBLAS/cuFFT/NCCL/cuDNN/cuSOLVER in the library: " + (", ".join(libs) if libs else "none"))


if __name__ == "__main__":
    sys.exit(main())
