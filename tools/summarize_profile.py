#!/usr/bin/env python
"""Turn the ncu outputs of a gpurun call into the tracked summary under profiles/.

    python tools/summarize_profile.py --launches gpurun_out/launches.csv --last N \
        --rep gpurun_out/prof.ncu-rep --title "..." --out profiles/<name>.md

--launches : csv of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`
--last     : keep only the last N launches (= the profiled step; earlier ones are set-up / warm-up)
--rep      : one or more `ncu --set full` reports; their raw page is read with `ncu -i`
"""
import argparse
import collections
import csv
import io
import re
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").strip()


def launches_table(path, last):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
            rows.append((short(r["Kernel Name"]), ns))
    if last:
        rows = rows[-last:]
    agg = collections.OrderedDict()
    for k, ns in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    out = [f"Total kernel time of the step under ncu: {tot / 1e6:.1f} ms over {len(rows)} launches", "",
           "| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {ns / 1e6:.3f} | {ns / tot:.3f} |")
    return "\n".join(out)


def rep_table(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = [f"`{short(data[0][hdr.index('Kernel Name')])}` — {len(data)} captured launch(es), `{path}`", "",
           "| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(data))) + " |",
           "|---|" + "---:|" * len(data)]
    for m, label in METRICS:
        if m in hdr:
            i = hdr.index(m)
            out.append(f"| {label} ({units[i]}) | " + " | ".join(d[i] for d in data) + " |")
    return "\n".join(out)


def traffic_entry(path):
    """dram bytes (read + write) and duration per captured launch of one ncu --set full report."""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    num = lambda v: float(v.replace(",", ""))
    tot = [num(d[ir]) * scale[units[ir]] + num(d[iw]) * scale[units[iw]] for d in data]
    dur = [num(d[it]) * tscale[units[it]] for d in data]
    return short(data[0][hdr.index("Kernel Name")]), {"dram_bytes_per_launch": sum(tot) / len(tot), "duration_us_under_ncu": sum(dur) / len(dur),
                                                       "launches_captured": len(data), "report": path}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traffic", help="also write {kernel: dram bytes per launch} of the --rep reports to this json (bench.py reads it)")
    ap.add_argument("--launches")
    ap.add_argument("--last", type=int, default=0)
    ap.add_argument("--rep", nargs="*", default=[])
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--note", default="")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    parts = [f"# {a.title}", ""]
    if a.note:
        parts += [a.note, ""]
    if a.launches:
        parts += [launches_table(a.launches, a.last), ""]
    for r in a.rep:
        parts += [rep_table(r), ""]
    if a.traffic:
        import json
        kernels = dict(traffic_entry(r) for r in a.rep)
        with open(a.traffic, "w") as f:
            json.dump({"source": a.out, "metric": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full --clock-control none)",
                       "kernels": kernels}, f, indent=1)
    with open(a.out, "w") as f:
        f.write("\n".join(parts))
    print("\n".join(parts))


if __name__ == "__main__":
    main()
