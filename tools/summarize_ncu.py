"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r3.csv profiles/launches_r3.md "title"
  python tools/summarize_ncu.py full     gpurun_out/prof_r3.ncu-rep profiles/ncu_full_r3.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        key = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, "w") as f:
        f.write("# %s\n\nncu `--metrics gpu__time_duration.sum --clock-control none` over ONE eager forward "
                "(cold-cache, serialised launches: compare shares, not absolutes).\nSource: `%s`.\n\n" % (title, src))
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% | %.1f |\n" % (k, n, t / 1e3, 100 * t / tot, t / n))
        f.write("| **total** | %d | %.3f | 100%% | |\n" % (sum(n for n, _ in agg.values()), tot / 1e3))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    # machine-readable per-kernel DRAM traffic (bench.py's roofline.traffic reads this file)
    import json
    traffic = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        def val(k):
            v = float(d[k].replace(",", ""))
            u = units[hdr.index(k)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        name = re.sub(r"\(.*", "", d.get("Kernel Name", "?"))
        e = traffic.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        tu = units[hdr.index("gpu__time_duration.sum")].lower()
        e["time_us"] += float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(tu, 1)
    with open(dst.replace(".md", ".json"), "w") as f:
        json.dump({"source": src, "kernels": traffic}, f, indent=1)
    with open(dst, "w") as f:
        f.write("# ncu --set full summary\n\nSource: `%s` (`ncu --set full --clock-control none --import-source on`).\n\n" % src)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("## %s  (grid %s, block %s)\n\n| metric | value | unit |\n|---|---:|---|\n" % (
                re.sub(r"\(.*", "", d.get("Kernel Name", "?"))[:90], d.get("Grid Size", "?"), d.get("Block Size", "?")))
            for k in hdr:
                if any(k == kk or k.startswith(kk) for kk in KEYS):
                    f.write("| %s | %s | %s |\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "kernel launch list")
    else:
        full(sys.argv[2], sys.argv[3])
