"""Instruction histogram per kernel of libuwm_b200.so (cuobjdump -sass), written to profiles/sass_digest.txt:

    python tools/sass_digest.py

Evidence that the shipped library is hand-written tcgen05 / TMEM / TMA code: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld),
UTMALDG / UTMASTG (TMA tensor loads / stores), UTCBAR (tcgen05.commit), SYNCS (mbarrier), per kernel template.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unet_watermark_b200", "lib", "libuwm_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "LDGSTS", "ACQBULK", "UTMAPF", "HMMA", "FFMA",
        "LDG", "STG", "LDS", "STS", "ATOM", "RED", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()  # noqa: E731
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            base = op.split(".")[0]
            kernels[cur][base] += 1
            if op.startswith("UTCHMMA.2CTA"):
                kernels[cur]["UTCHMMA.2CTA"] += 1
    lines = [f"SASS digest of {os.path.relpath(LIB, ROOT)} ({len(kernels)} kernels; cuobjdump -sass, sm_100a)", ""]
    tot = collections.Counter()
    hdr = f"{'kernel':<92s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in KEYS[:10])
    lines.append(hdr)
    for name, c in kernels.items():
        d = demangle(name)
        d = d.replace("(anonymous namespace)::", "")                  # before the argument list is cut at its '('
        d = re.sub(r"\(.*", "", d).replace("void ", "").replace("uwm::", "")
        d = d.replace("(int)", "").replace("(bool)", "")
        lines.append(f"{d[:92]:<92s} {c['_total']:>7d} " + " ".join(f"{c[k]:>8d}" for k in KEYS[:10]))
        tot.update(c)
    lines += ["", "totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]), f"all instructions: {tot['_total']}"]
    path = os.path.join(ROOT, "profiles", "sass_digest.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-3:]))
    print(path)


if __name__ == "__main__":
    sys.exit(main())
