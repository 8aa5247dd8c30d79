"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output: top stall sites."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    si, so, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    data = []
    for i, r in enumerate(rows[hi + 1:]):
        if len(r) <= si:
            continue
        try:
            data.append((int(r[si] or 0), r[so], int(r[ie] or 0), i))
        except ValueError:
            continue
    tot = sum(d[0] for d in data)
    print('total samples', tot, 'instructions', len(data), 'warp-instr executed', sum(d[2] for d in data))
    for d in sorted(data, reverse=True)[:top]:
        print('%7d %5.1f%%  #%4d exec=%9d  %s' % (d[0], 100.0 * d[0] / max(tot, 1), d[3], d[2], d[1][:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
