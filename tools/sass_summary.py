"""Static evidence for profiles/: per kernel of libltephy_b200.so the register / shared-memory figures ptxas reported (build.log)
and the SASS mnemonic histogram (cuobjdump -sass), with the instructions that matter for this design called out:
VIADD.16x2 / VIMNMX.S16x2 / VIADDMNMX.S16x2 (two trellises per register in the Viterbi and turbo kernels), UBLKCP (cp.async.bulk
staging in the OFDM kernel), SHFL / VOTE (warp exchange), LDS/STS vs LDG/STG balance.  usage: python tools/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ltesniffer_b200", "libltephy_b200.so")
KEY = ["VIADD.16x2", "VIMNMX.S16x2", "VIADDMNMX.S16x2", "UBLKCP", "SYNCS", "SHFL", "VOTE", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ATOMS", "BAR", "FADD", "FMUL", "FFMA", "IMAD", "LOP3", "PRMT"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    regs = {}
    log = open(os.path.join(ROOT, "ltesniffer_b200", "build.log")).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", log, re.S):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3) or 0), int(m.group(4) or 0))
    print("# SASS summary of", os.path.relpath(SO, ROOT), "(sm_100a); counts are static instruction counts")
    for k, c in kernels.items():
        total = sum(c.values())
        r = regs.get(k)
        name = demangle(k)
        name = re.sub(r"\(.*", "", name)
        print("\n== %s   %d SASS instructions%s" % (name, total, "   registers %d, stack %d B, static smem %d B" % r if r else ""))
        fam = collections.Counter()
        for op, n in c.items():
            for key in KEY:
                if op == key or op.startswith(key + ".") or (key in ("VIADD.16x2", "VIMNMX.S16x2", "VIADDMNMX.S16x2") and op.startswith(key)):
                    fam[key] += n
                    break
        print("   " + "  ".join("%s %d" % (k2, fam[k2]) for k2 in KEY if fam[k2]))
        top = ", ".join("%s %d" % (op, n) for op, n in c.most_common(8))
        print("   top: " + top)


if __name__ == "__main__":
    sys.exit(main())
