"""Opcode summary of the kernels in libvalunc.so (developer tool; evidence for profiles/):
    python bench/sass_summary.py > profiles/<tag>_sass_opcodes.txt
Per kernel: instructions, and the counts of the mnemonics that show what the code is built on -- UBLKCP (cp.async.bulk, the
1-D TMA engine copy), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (Blackwell packed fp32), MUFU.LG2 / EX2 / RCP, REDUX, MATCH,
SHFL, VOTE, LDS / STS, LDG / STG, ATOM / REDG (global reductions), LDL / STL (spills) -- and whether any tensor-core opcode appears (none should:
there is no contraction on this path)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffuncertainty_b200", "lib", "libvalunc.so")
WANT = ["UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "MUFU.LG2", "MUFU.EX2", "MUFU.RCP", "REDUX", "MATCH", "SHFL", "VOTE",
        "LDS", "STS", "LDG", "STG", "ATOM", "REDG", "LDL", "STL", "UTCMMA", "HMMA", "LDTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    name, counts, total, arch = None, None, 0, set()
    rows = []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, total, counts))
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            total += 1
            for w in WANT:
                if op == w or op.startswith(w + ".") or (w.startswith("MUFU") and op.startswith(w)):
                    counts[w] += 1
    if name:
        rows.append((name, total, counts))
    demangled = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print("cubin architectures:", ", ".join(sorted(arch)))
    print(f"{len(rows)} kernels")
    agg = collections.Counter()
    for (n, total, c), d in zip(rows, demangled):
        short = re.sub(r"\(.*", "", d.replace("vu::", ""))[:90]
        agg.update(c)
        print(f"{short:92s} {total:6d} instr  " + " ".join(f"{k}={v}" for k, v in c.items() if v))
    print("\nall kernels: " + " ".join(f"{k}={agg[k]}" for k in WANT))
    print("tensor-core opcodes (UTCMMA / HMMA / LDTM):", agg["UTCMMA"] + agg["HMMA"] + agg["LDTM"], "-- none by design (no contraction on this path)")


if __name__ == "__main__":
    main()
