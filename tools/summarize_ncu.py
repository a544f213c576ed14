"""Turn gpurun_out/*.csv / *.ncu-rep into the small text summaries committed under profiles/."""
import collections
import csv
import subprocess
import sys


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"])
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        agg.setdefault(row["Kernel Name"][:90], []).append(v)
    # kernels that only the one-off validation of the step launches (bench.validate_step: the fp32 FFMA deform-conv path the
    # tensor-core results are compared with, torch reductions of the comparisons) and the eager end-to-end arm's torch casts
    val_only = ("dcn_fwd_kernel(", "dcn_bwd_data_kernel(", "dcn_bwd_weight_kernel(", "reduce_kernel", "AbsFunctor", "direct_copy_kernel",
                "bfloat16_copy_kernel", "CUDAFunctor_add<c10::BFloat16>", "MaxNanFunctor", "index_elementwise", "CompareFunctor")
    step = collections.OrderedDict((k, v) for k, v in agg.items() if not any(t in k for t in val_only))

    def table(f, rows, title):
        tot = sum(sum(v) for v in rows.values())
        f.write("\n## %s\n%-92s %5s %10s %7s\n" % (title, "kernel", "n", "avg_us", "share"))
        for k, v in sorted(rows.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-92s %5d %10.1f %6.1f%%\n" % (k, len(v), sum(v) / len(v), 100 * sum(v) / tot))

    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 2 --warmup 1`\n"
                "# (cold-cache, serialised launches: compare SHARES, not absolute times)\n")
        table(f, step, "kernels of the step (device-resident graphs, per-stage graphs, end-to-end arms), validation-only kernels excluded")
        table(f, agg, "every launch of the command, incl. the one-off validation against the fp32 FFMA path / the oracle")
    print(open(out).read())


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on ; per-launch raw metrics\n")
        for r in rows[2:]:
            f.write("\n## %s (launch id %s)\n" % (r[idx["Kernel Name"]][:100], r[idx["ID"]]))
            for w in want:
                if w in idx:
                    f.write("%-72s %18s %s\n" % (w, r[idx[w]], units[idx[w]]))
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3])
