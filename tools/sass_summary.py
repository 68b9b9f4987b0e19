"""SASS opcode summary per kernel of libmquant_b200.so (cuobjdump -sass; no GPU needed): the
mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md) plus the instruction total.
python tools/sass_summary.py > profiles/<round>_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mquant_b200", "libmquant_b200.so")
KEY = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF",
       "SYNCS", "HMMA", "IMMA", "IDP", "MUFU", "REDUX", "ATOMG", "RED", "LDGSTS"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            total[name] += 1
            op = m.group(1)
            for k in KEY:
                if op == k or op.startswith(k + "."):
                    counts[name][k] += 1
    print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (sm_100a); only the listed mnemonics + total")
    for n in sorted(counts, key=lambda k: -total[k]):
        c = counts[n]
        print(f"{n}\n    total {total[n]}  " + "  ".join(f"{k} {c[k]}" for k in KEY if c[k]))


if __name__ == "__main__":
    main()
