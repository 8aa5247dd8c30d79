"""Per-kernel SASS evidence of libbcg_b200.so (run here, no GPU needed):
    python profiles/sass_summary.py > profiles/sass_summary.txt
Counts the mnemonics B200_PROFILING.md names: UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (bulk copy engine),
LDGSTS (cp.async), UTC*MMA / LDTM / STTM (tcgen05 -- none expected: nothing in this path is a contraction), plus the
warp-collective and fp64 instructions that characterise the step kernels."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bc_gym_planning_env_b200", "csrc", "libbcg_b200.so")
OPS = ["UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "UTC", "LDTM", "STTM", "HMMA", "VOTE", "SHFL", "ATOM", "ATOMS", "RED", "DFMA", "DMUL", "DADD", "MUFU", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = re.findall(r"code for (sm_\w+)", out)
    print("# %s: %s" % (os.path.relpath(LIB, ROOT), ", ".join(sorted(set(arch)))))
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
            kernels[name] = collections.Counter()
            continue
        m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            kernels[name]["total"] += 1
            for o in OPS:
                if op.startswith(o) and (o != "ATOM" or not op.startswith("ATOMS")):
                    kernels[name][o] += 1
    print("%-46s %6s  %s" % ("kernel", "SASS", "  ".join("%s" % o for o in OPS)))
    for k, c in kernels.items():
        print("%-46s %6d  %s" % (k[:46], c["total"], "  ".join("%*d" % (len(o), c[o]) for o in OPS)))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("\nTMA tensor loads (UTMALDG): %d, bulk-copy engine (UBLKCP): %d, cp.async (LDGSTS): %d, tcgen05 (UTC*MMA / LDTM / STTM): %d "
          "(none: no contraction in this path), legacy tensor path (HMMA): %d" % (tot["UTMALDG"], tot["UBLKCP"], tot["LDGSTS"],
                                                                            tot["UTC"] + tot["LDTM"] + tot["STTM"], tot["HMMA"]))


if __name__ == "__main__":
    main()
