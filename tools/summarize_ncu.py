"""Summarise ncu outputs into profiles/ (text committed to git).
usage: python tools/summarize_ncu.py <launches.csv> <full.ncu-rep> <out.md> [title]"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "sm__cycles_active.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][-70:]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["| launches | total us | share | kernel |", "|---:|---:|---:|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        out.append(f"| {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | `{k}` |")
    out.append(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
    return "\n".join(out)


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return "(no data)"
    h, u = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = r[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        out.append(f"### `{name[:90]}`\n\n| metric | unit | value |\n|---|---|---:|")
        for i, n in enumerate(h):
            if n in KEYS:
                out.append(f"| {n} | {u[i]} | {r[i]} |")
    return "\n".join(out)


if __name__ == "__main__":
    lst, rep, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else "ncu summary"
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write(launches(lst) + "\n\n## Top kernel (`ncu --set full --clock-control none`)\n\n" + full(rep) + "\n")
    print(open(dst).read()[:3000])
