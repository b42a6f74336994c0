#!/usr/bin/env python
"""Turn the scratch artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py <tag> <bench.json> <launches.csv> [<full.ncu-rep>]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {100 * v[1] / tot:.1f}% |")
    return "\n".join(out)


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"), ("lts__t_bytes.sum", "L2 bytes"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
    out = ["| kernel | " + " | ".join(w[1] for w in want) + " |", "|---|" + "---|" * len(want)]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for m, _ in want:
            cells.append(f"{r[idx[m]]} {units[idx[m]]}".strip() if m in idx else "-")
        out.append(f"| {name} | " + " | ".join(cells) + " |")
    return "\n".join(out)


def main():
    tag, bench, lcsv = sys.argv[1:4]
    rep = sys.argv[4] if len(sys.argv) > 4 else None
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    b = json.load(open(bench))
    json.dump(b, open(os.path.join(ROOT, "profiles", f"{tag}_bench.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
        f.write(open(lcsv).read())
    md = [f"# {tag}: ncu summaries (B200, 640x480 stream, `bench.py`)", "",
          f"bench line: value {b['value']:.1f} fps ({b['ms_per_step']:.4f} ms/step), e2e {b['e2e']['value']:.1f} fps, "
          f"cpu_baseline {b.get('cpu_baseline', {}).get('value', float('nan')):.1f} fps ({b.get('cpu_baseline', {}).get('cores', '?')} cores), "
          f"gpu_launches {b['gpu_launches']}", "",
          "## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
          launches(lcsv), ""]
    if rep:
        md += ["## `ncu --set full --clock-control none` (one launch per kernel; ncu flushes caches between replays, so",
               "per-kernel DRAM reads of data that is L2-resident inside a real step show up as DRAM traffic here)", "", full(rep), ""]
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(md))
    print("\n".join(md))


if __name__ == "__main__":
    main()
