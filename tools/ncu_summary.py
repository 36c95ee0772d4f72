"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread ', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum ', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled', 'launch__grid_size', 'launch__block_size', 'smsp__average_warp_latency_per_inst_issued',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum ', 'l1tex__t_bytes.sum ']


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")][:90])
        for h, u, v in zip(hdr, units, vals):
            if any((h + ' ').startswith(k) or k in h for k in KEYS):
                try:
                    if float(v.replace(',', '')) == 0 and 'stalled' in h:
                        continue
                except ValueError:
                    pass
                print("  %-88s %-12s %s" % (h, u, v))


if __name__ == "__main__":
    main(sys.argv[1])
