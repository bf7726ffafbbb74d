"""Per-kernel SASS evidence of the shipped library (profiles/sass_rNN.txt): counts of the mnemonics that prove
Blackwell-native code paths (B200_PROFILING.md) in every kernel of libvitb200.so.

    python tools/sass_evidence.py > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vit.triton_b200", "vit", "kernels", "libvitb200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCHMMA.2CTA", "UTCQMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG*MULTICAST", "UTMASTG", "UTCBAR", "MUFU.EX2",
             "FFMA2", "HMMA", "LDGSTS", "RED.E", "ATOMG", "ST.E.STRONG.SYS", "LDG.E.STRONG.SYS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("vt::", "").replace("void ", "")
            depth, cut = 0, len(name)
            for i, ch in enumerate(name):          # cut the parameter list: first "(" outside template brackets
                if ch == "<":
                    depth += 1
                elif ch == ">":
                    depth -= 1
                elif ch == "(" and depth == 0:
                    cut = i
                    break
            name = name[:cut]
            cur = kernels.setdefault(name, collections.Counter())
            cur["_variants"] += 1
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_instructions"] += 1
        if op.startswith("UTMALDG") and ".MULTICAST" in op:
            cur["UTMALDG*MULTICAST"] += 1
        for mn in MNEMONICS:
            if op == mn or (op.startswith(mn + ".") and not (mn in ("UTCHMMA", "UTCQMMA") and ".2CTA" in op)):
                cur[mn] += 1
    print(f"# SASS evidence for {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes), cuobjdump -sass, sm_100a")
    print("# columns: instantiations | instructions | " + " | ".join(MNEMONICS))
    totals = collections.Counter()
    for name, c in sorted(kernels.items(), key=lambda kv: -kv[1]["_instructions"]):
        print(f"{name}: {c['_variants']} | {c['_instructions']} | " + " | ".join(str(c[m]) for m in MNEMONICS))
        totals.update(c)
    print("TOTAL: " + " | ".join(f"{m}={totals[m]}" for m in MNEMONICS))


if __name__ == "__main__":
    sys.exit(main())
