"""Table of selected metrics per launch from an `ncu --page raw --csv` export."""
import csv
import sys

KEYS = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("gpu__time_duration.sum", "ms"),
        ("smsp__inst_executed.sum", "inst"), ("sm__inst_executed.avg.per_cycle_active", "ipc"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit"), ("lts__t_sector_hit_rate.pct", "l2hit"),
        ("sm__icc_request_hit_rate", "icc_hit"), ("gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "gcc%"),
        ("smsp__average_warp_latency_per_inst_issued.ratio", "cyc/inst"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "short_sb"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "no_inst"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "wait"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "branch"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "lg_thr"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "membar"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "sleep"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "dispatch"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "math"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "mio"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "not_sel"),
        ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "selected"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = []
    for k, short in KEYS:
        m = [i for i, h in enumerate(hdr) if h == k or h.startswith(k)]
        idx.append((short, m[0] if m else None))
    for vals in rows[2:]:
        out = []
        for short, i in idx:
            if i is None:
                continue
            v = vals[i]
            if short == "kernel":
                v = v.split("(")[0][-22:]
            else:
                try:
                    f = float(v.replace(",", ""))
                    v = ("%.3g" % f) + (units[i][:2] if short in ("ms", "dram_rd", "dram_wr") else "")
                except ValueError:
                    pass
            out.append("%s=%s" % (short, v))
        print(" ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
