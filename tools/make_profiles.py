"""profiles/ from gpurun_out/: launch list, per-kernel DRAM traffic of one env-step, full ncu text of the
dominant kernel (+ the instruction-cache metrics), GoL kernel summary, bench line."""
import collections
import csv
import json
import shutil
import subprocess
import sys

R = "/root/repo/"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def full_text(rep, dst, keys):
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    h, u, vals = raw(rep)
    with open(dst, "w") as f:
        f.write(det)
        f.write("\n==== selected raw metrics ====\n")
        for n, un, v in zip(h, u, vals[0]):
            if any(k in n for k in keys) and "TriageCompute" not in n:
                f.write("  %-90s %-8s %s\n" % (n, un, v))


def traffic():
    src = R + "profiles/r01_launches.csv"
    all_rows = list(csv.reader(open(src)))
    hdr = [r for r in all_rows if r and r[0] == "ID"][0]
    iu = hdr.index("Metric Unit")
    by = collections.OrderedDict()
    for r in all_rows:
        if len(r) <= 12 or not r[0].isdigit():
            continue
        v, u = float(r[-1].replace(",", "")), r[iu]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        by.setdefault(r[0], {"name": r[4]})[r[-3]] = v
    ids = list(by)
    sc = [i for i, k in enumerate(ids) if "k_mp_sched_scatter" in by[k]["name"]]
    lo, hi = sc[-4] + 1, sc[-3] + 1
    agg = collections.OrderedDict()
    for k in ids[lo:hi]:
        d = by[k]
        n = d["name"].split("(")[0].split("::")[-1]
        a = agg.setdefault(n, {"launches": 0, "ms_serialised": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0})
        a["launches"] += 1
        a["ms_serialised"] += d["gpu__time_duration.sum"]
        a["dram_read_MB"] += d["dram__bytes_read.sum"] / 1e6
        a["dram_write_MB"] += d["dram__bytes_write.sum"] / 1e6
    for a in agg.values():
        for k in a:
            a[k] = round(a[k], 3)
    tot = sum(a["dram_read_MB"] + a["dram_write_MB"] for a in agg.values()) * 1e6
    tms = sum(a["ms_serialised"] for a in agg.values())
    out = {"source": "profiles/r01_launches.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none; python bench.py --steps 2 --warmup 3 --age-steps 10 --no-cpu-baseline; one env-step of "
                     "32768 envs = launches %d..%d: 3 env groups, 50 frames per launch (100 for the heavy group), 32 envs per "
                     "CTA; per-launch times are cold-cache and serialised)" % (lo, hi - 1),
           "dram_bytes_per_step_32768_envs": tot, "dram_bytes_per_env_tick": tot / 32768,
           "algorithmic_bytes_per_env_tick": 161533,
           "share_of_step_time": {k: round(a["ms_serialised"] / tms, 4) for k, a in agg.items()}, "kernels": agg}
    json.dump(out, open(R + "profiles/r01_traffic.json", "w"), indent=1)
    print(json.dumps(out["share_of_step_time"]), out["dram_bytes_per_env_tick"])


if __name__ == "__main__":
    shutil.copy(R + "gpurun_out/r01_launches.csv", R + "profiles/r01_launches.csv")
    shutil.copy(R + "gpurun_out/bench_r01.json", R + "profiles/r01_bench_line.json")
    traffic()
    full_text(R + "gpurun_out/r01_k_mp_frames.ncu-rep", R + "profiles/r01_k_mp_frames_ncu_full.txt",
              ("gcc__", "sm__icc", "idc__", "smsp__issue_inst0", "smsp__issue_active.avg.pct", "dram__bytes_read.sum",
               "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst"))
    full_text(R + "gpurun_out/r01_k_gol_step.ncu-rep", R + "profiles/r01_k_gol_step_ncu_full.txt",
              ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "dram__throughput", "smsp__inst_executed.sum"))
