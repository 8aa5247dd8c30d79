"""Turn an .ncu-rep into the small text summary committed under profiles/:
    python profiles/summarise_ncu.py gpurun_out/prof.ncu-rep > profiles/rN_summary.txt
Needs `ncu` on PATH (no GPU needed to read a report)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_instr"),
    ("sm__warps_active.avg.per_cycle_active", "warps_per_sm"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs_blocks"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem_blocks"),
    ("smsp__inst_executed_op_tma_ld.sum", "tma_ld_instr"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "tma_bytes"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall_long_scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall_barrier"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall_wait"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall_short_scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "stall_selected(issuing)"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall_not_selected"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    print("# %s (ncu --set full --clock-control none; per launch, cold-ish cache, serialised)" % path)
    for r in rows[2:]:
        print("\n%s  grid=%s block=%s" % (r[h.index("Kernel Name")].split("(")[0], r[h.index("Grid Size")], r[h.index("Block Size")]))
        for key, label in KEYS:
            if key in h and r[h.index(key)] not in ("", "n/a"):
                print("    %-28s %s %s" % (label, r[h.index(key)], units[h.index(key)]))


if __name__ == "__main__":
    main(sys.argv[1])
