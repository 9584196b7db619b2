"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md):
tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), LDGSTS (cp.async), legacy tensor cores
(HMMA), packed fp32 (FADD2 / FFMA2).  Reads the built library with cuobjdump; writes profiles/sass_summary.txt.
    python tools/sass_summary.py [libwhisper_b200.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "whisper-rust-ort_b200", "libwhisper_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "FADD2", "FFMA2", "SYNCS", "UCGABAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur, n_inst = {}, [], None, collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if cur and m:
            op = m.group(1)
            n_inst[cur] += 1
            for k in MNEMONICS:
                if op.startswith(k):
                    counts[cur][k] += 1
    names = demangle(order)
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass; sm_100a)")
    print(f"# {'kernel':<100} {'instr':>6} " + " ".join(f"{k:>8}" for k in MNEMONICS))
    for f in order:
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", names[f])
        short = re.sub(r"\(.*", "", short)
        print(f"{short[:100]:<102} {n_inst[f]:>6} " + " ".join(f"{counts[f][k]:>8}" for k in MNEMONICS))


if __name__ == "__main__":
    main()
