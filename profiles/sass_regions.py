"""Where a kernel's issued instructions and stall samples go, in chunks of N SASS instructions:
    ncu -i X.ncu-rep --page source --csv [--kernel-name regex:K] > src.csv ; python profiles/sass_regions.py src.csv [N]
Each line: share of warp instructions executed, share of stall samples, and the memory / sync opcodes in the chunk."""
import csv
import sys
from collections import Counter

MARK = ('LDG', 'ATOMS', 'STS', 'LDS', 'BAR', 'SHFL', 'UBLKCP', 'POPC', 'FLO', 'F2I', 'DMUL', 'DFMA', 'DADD', 'VOTE', 'STG', 'LDGSTS')


def main(path, step=40):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    si, so, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    data = []
    for r in rows[hi + 1:]:
        if len(r) > si:
            try:
                data.append((int(r[si] or 0), r[so].strip(), int(r[ie] or 0)))
            except ValueError:
                pass
    tot_i, tot_s = sum(d[2] for d in data), sum(d[0] for d in data)
    print("%d SASS instructions, %d warp instructions executed, %d samples" % (len(data), tot_i, tot_s))
    for c in range(0, len(data), step):
        ch = data[c:c + step]
        ops = [d[1].split()[1] if d[1].startswith('@') else d[1].split()[0] for d in ch]
        marks = Counter(o.split('.')[0] for o in ops if any(m in o for m in MARK))
        print('%4d-%4d instr %5.1f%% samples %5.1f%%  %s' % (c, c + step - 1, 100. * sum(d[2] for d in ch) / tot_i,
                                                            100. * sum(d[0] for d in ch) / max(tot_s, 1), dict(marks)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
