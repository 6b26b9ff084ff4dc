"""DRAM traffic of the flow kernels from an `ncu --csv` launch list with dram__bytes_read/write and gpu__time_duration
(tools/ncu_launches.py prints the per-kernel table): bytes per window and the kernels' share of the step.

    python tools/ncu_traffic.py gpurun_out/launches_r02.csv WINDOWS_TOTAL > profiles/traffic_r02.json
"""
import collections
import csv
import json
import sys


def main(path, windows):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        per.setdefault((r["ID"], r["Kernel Name"]), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "") or 0)
    flow = [(k[1], m) for k, m in per.items() if "hgsfa::" in k[1]]
    layer = [m for name, m in flow if "front_kernel" in name or "layer_" in name or "back_kernel" in name]
    t_all = sum(m["gpu__time_duration.sum"] for _, m in flow)
    by_kernel = collections.OrderedDict()
    for name, m in flow:
        short = name.split("hgsfa::")[1].split("(")[0].split("<")[0]
        a = by_kernel.setdefault(short, dict(launches=0, time_ns=0.0, dram_bytes=0.0))
        a["launches"] += 1
        a["time_ns"] += m["gpu__time_duration.sum"]
        a["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    for a in by_kernel.values():
        a["dram_bytes_per_window"] = a["dram_bytes"] / windows
        a["share_of_flow_time"] = a["time_ns"] / t_all
    b = sum(m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0) for m in layer)
    print(json.dumps({
        "windows": windows, "layer_launches": len(layer), "layer_dram_bytes": b, "layer_dram_bytes_per_window": b / windows,
        "layer_time_ns_ncu": sum(m["gpu__time_duration.sum"] for m in layer),
        "layer_share_of_step": sum(m["gpu__time_duration.sum"] for m in layer) / t_all, "kernels": by_kernel,
        "source": "%s (ncu dram__bytes_read.sum + dram__bytes_write.sum, %d windows, fused front + layer launches)" % (path.replace("gpurun_out", "profiles"), windows)},
        indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
