"""Turn ncu exports into the committed summaries.

  python profiles/summarize_ncu.py raw   <raw.csv>  <out.md> [traffic.json]   # from: ncu -i X.ncu-rep --page raw --csv
  python profiles/summarize_ncu.py list  <launches.csv> <out.md>              # from: ncu --metrics gpu__time_duration.sum --csv

`traffic.json` maps bench.py's kernel classes to dram__bytes_read.sum + dram__bytes_write.sum per launch of the
finest-level instance (the largest launch of each kernel), which bench.py reports as roofline.traffic.
"""
import collections
import csv
import json
import re
import sys

CLASS_OF = [
    (r"stencil_march_kernel<(\(int\))?1", "apply_dot"), (r"stencil_march_kernel<(\(int\))?4", "cheb_zero"),
    (r"stencil_march_kernel<(\(int\))?2", "residual|cheb_first"), (r"stencil_march_kernel<(\(int\))?3", "cheb_next"),
    (r"stencil_march_kernel<(\(int\))?5", "cheb_next"), (r"prolong_add3d", "prolong_add"), (r"restrict_kernel", "restrict"),
    (r"restrict3d_kernel", "restrict"), (r"xp_update_kernel", "xp_update"), (r"r_update_kernel", "r_update"),
    (r"axpy2_kernel", "axpy2"), (r"aypx_dev_kernel", "aypx"), (r"dot2_kernel", "dot2"),
]


def classify(name):
    for pat, cls in CLASS_OF:
        if re.search(pat, name):
            return cls
    return None


def raw(path, out, traffic_path=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
    want = [w for w in want if w in ix]

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    def to_us(v, u):
        v = float(v.replace(",", ""))
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)

    best = {}
    lines = ["| kernel | grid | block | regs | time us | DRAM read MB | DRAM write MB | DRAM % | SM % | issue % | L2 hit % |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]]
        t = to_us(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        short = re.sub(r"\(.*", "", name.replace("void p4b::", "").replace("p4b::", ""))[:48]
        g = lambda k: r[ix[k]] if k in ix else "-"
        lines.append("| %s | %s | %s | %s | %.1f | %.1f | %.1f | %s | %s | %s | %s |" % (
            short, r[ix["Grid Size"]], r[ix["Block Size"]], g("launch__registers_per_thread"), t, rd / 1e6, wr / 1e6,
            g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            g("smsp__issue_active.avg.pct"), g("lts__t_sector_hit_rate.pct")))
        cls = classify(name)
        if cls and (cls not in best or t > best[cls][0]):
            best[cls] = (t, rd + wr)
    open(out, "w").write("\n".join(lines) + "\n")
    if traffic_path:
        traffic = {}
        for cls, (t, b) in best.items():
            for c in cls.split("|"):
                traffic[c] = b
        json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
        print(json.dumps(traffic, indent=1, sort_keys=True))


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]].replace("void p4b::", ""))[:60]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ix["Metric Unit"]], 1)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["total device time of the listed launches: %.1f us (cold-cache, serialised: compare SHARES)" % tot, "",
             "| kernel | launches | us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.1f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:16]))


if __name__ == "__main__":
    if sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        launch_list(sys.argv[2], sys.argv[3])
