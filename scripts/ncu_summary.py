#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: a launch list CSV (--metrics gpu__time_duration.sum) and/or a
.ncu-rep full capture (read with `ncu -i ... --page raw --csv`).
`--traffic KEY=file.ncu-rep[:kernel substring]` records dram__bytes_read.sum + dram__bytes_write.sum of the (first
matching) kernel in profiles/ncu_traffic.json, which bench.py reports as `roofline.traffic`."""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum"]


def launches(path):
    text = open(path).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r["Kernel Name"].split("(")[0], []).append(float(r["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    print(f"## launch list: {path}  ({len(rows)} launches, gpu__time_duration.sum, cold-cache serialised)")
    print("| kernel | launches | avg us | share of listed time |\n|---|---|---|---|")
    for k, v in agg.items():
        print(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/tot:.3f} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"## full capture: {path}")
    for r in rows[2:]:
        print(f"\n### `{r[hdr.index('Kernel Name')]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")


def traffic(arg):
    import json
    import os

    key, rest = arg.split("=", 1)
    path, _, filt = rest.partition(":")
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if filt in name:
            rd, wr = (int(float(r[hdr.index(k)])) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            dur = float(r[hdr.index("gpu__time_duration.sum")]) / 1e3
            dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json")
            d = json.load(open(dst)) if os.path.exists(dst) else {}
            d[key] = {"bytes": rd + wr, "read": rd, "write": wr, "kernel": name.split("(")[0][:80], "ncu_us": round(dur, 2),
                      "source": f"ncu --set full, {os.path.basename(path)}: dram read+write of one launch"}
            json.dump(d, open(dst, "w"), indent=1, sort_keys=True)
            print(f"{key}: {rd + wr} B ({name[:60]}, {dur:.1f} us)")
            return
    print(f"{key}: no kernel matching {filt!r} in {path}")


if __name__ == "__main__":
    args = sys.argv[1:]
    while args and args[0] == "--traffic":
        traffic(args[1])
        args = args[2:]
    for a in args:
        (full if a.endswith(".ncu-rep") else launches)(a)
        print()
