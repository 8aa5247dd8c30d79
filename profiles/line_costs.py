"""Warp instructions executed (and stall samples) per CUDA source line of one kernel:
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > cs.csv
    python profiles/line_costs.py cs.csv [units] [min_share_percent]
`units` divides the counts (e.g. 65536 envs) so that a line reads as warp instructions per env."""
import csv
import sys


def main(path, units=1.0, min_share=0.3):
    cur_file, cur_line, cur_src = None, None, ""
    cost = {}
    col = None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            col = (r.index("Instructions Executed"), r.index("# Samples"))
            continue
        if col is None or len(r) <= col[0]:
            continue
        if r[0] != "":
            cur_line, cur_src = int(r[0]), r[1].strip()
            continue
        try:
            ie, sm = int(r[col[0]]), int(r[col[1]] or 0)
        except ValueError:
            continue
        k = (cur_file, cur_line)
        c = cost.setdefault(k, [0, 0, 0, cur_src])
        c[0] += ie
        c[1] += sm
        c[2] += 1
    tot_i = sum(c[0] for c in cost.values())
    tot_s = sum(c[1] for c in cost.values())
    print("total %.1f warp instructions per unit, %d samples" % (tot_i / units, tot_s))
    for (f, ln), c in sorted(cost.items()):
        if 100.0 * c[0] / tot_i >= min_share or 100.0 * c[1] / max(tot_s, 1) >= min_share:
            print("%-16s %5d  instr %7.1f (%4.1f%%)  samples %4.1f%%  sass %3d | %s" % (
                f, ln, c[0] / units, 100.0 * c[0] / tot_i, 100.0 * c[1] / max(tot_s, 1), c[2], c[3][:110]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0, float(sys.argv[3]) if len(sys.argv) > 3 else 0.3)
