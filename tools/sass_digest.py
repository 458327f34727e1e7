#!/usr/bin/env python
"""Per-kernel digest of the SASS in libeamm_b200.so: counts of the tcgen05 / TMA / TMEM instructions that prove what
the hot path runs on (B200_PROFILING.md names the mnemonics).  Runs anywhere cuobjdump exists (no GPU needed):

    python tools/sass_digest.py > profiles/r2_sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "eamm_b200", "lib", "libeamm_b200.so")
PATTERNS = [
    ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"),          # tcgen05.mma kind::f16
    ("UTCQMMA", r"\bUTCQMMA\b(?!\.2CTA)"), ("UTCQMMA.2CTA", r"\bUTCQMMA\.2CTA"),          # tcgen05.mma kind::f8f6f4
    ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
    ("UTMAPF", r"\bUTMAPF|\bUTMACCTL"), ("SYNCS", r"\bSYNCS"), ("STG.256", r"\bSTG\.E\.ENL2\.256"),
    ("LDG.256", r"\bLDG\.E\.ENL2\.256"), ("F2FP.E4M3", r"F2FP\.SATFINITE\.E4M3"), ("F2FP.F16", r"F2FP\.SATFINITE\.F16"),
    ("HMMA/mma.sync", r"\bHMMA\b"), ("FFMA", r"\bFFMA\b"),
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        cur["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4,}\*/", line) else 0
        for name, pat in PATTERNS:
            if re.search(pat, line):
                cur[name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS digest of eamm_b200/lib/libeamm_b200.so (sm_100a), `python tools/sass_digest.py`")
    print("# columns: instruction count, then the non-zero counts of the mnemonics of interest")
    total = collections.Counter()
    for (mangled, cnt), nice in zip(kernels.items(), demangle):
        nice = re.sub(r"\(.*", "", nice)
        cols = "  ".join("%s=%d" % (n, cnt[n]) for n, _ in PATTERNS if cnt[n])
        print("%-72s %6d  %s" % (nice[:72], cnt["instructions"], cols))
        total.update(cnt)
    print("# total: " + "  ".join("%s=%d" % (n, total[n]) for n, _ in PATTERNS if total[n]))
    ptxas_resources()
    return 0


def ptxas_resources():
    """Registers / spills / static shared memory per kernel from the `-Xptxas -v` logs build.py keeps next to the objects."""
    libdir = os.path.dirname(LIB)
    logs = sorted(f for f in os.listdir(libdir) if f.endswith(".ptxas.log"))
    if not logs:
        return
    print("#\n# ptxas -v resource usage (eamm_b200/lib/*.ptxas.log): registers, barriers, static smem, stack, spill stores / loads")
    for f in logs:
        text = open(os.path.join(libdir, f)).read()
        entries = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                             r"(\d+) bytes spill loads\n.*?Used (\d+) registers, used (\d+) barriers(?:, \d+ bytes cumulative stack size)?"
                             r"(?:, (\d+) bytes smem)?", text)
        names = subprocess.run(["c++filt"], input="\n".join(e[0] for e in entries), capture_output=True, text=True).stdout.splitlines()
        for e, nice in zip(entries, names):
            nice = re.sub(r"\(.*", "", nice)
            print("%-72s regs=%-3s barriers=%s smem=%-6s stack=%-4s spill_st=%-4s spill_ld=%s"
                  % (nice[:72], e[4], e[5], e[6] or 0, e[1], e[2], e[3]))


if __name__ == "__main__":
    sys.exit(main())
