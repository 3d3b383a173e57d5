"""ncu `--page raw --csv` export -> profiles/*_kernel_traffic.json (DRAM bytes per launch and the headline
counters of every kernel in the capture; bench.py reads `dram_bytes` into `roofline.traffic`).

    python tools/ncu_traffic.py RAW.csv OUT.json --workload c3 --n-gpus 1 --storage lower --source "..."
"""
import argparse
import csv
import json
import re

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw"); ap.add_argument("out")
    ap.add_argument("--workload", default="c3"); ap.add_argument("--n-gpus", type=int, default=1)
    ap.add_argument("--storage", default="lower"); ap.add_argument("--source", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, default=None):
        i = col.get(name)
        if i is None or r[i] == "":
            return default
        try:
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        except ValueError:
            return default

    groups = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        name = re.sub(r"^(void )?(tal::)?", "", name)
        name = re.sub(r"\(.*$", "", name)
        groups.setdefault(name, []).append(r)
    out = {"source": a.source, "workload": a.workload, "n_gpus": a.n_gpus, "storage": a.storage, "kernels": {}}
    for name, rs in groups.items():
        def mean(metric):
            v = [val(r, metric) for r in rs]
            v = [x for x in v if x is not None]
            return sum(v) / len(v) if v else None
        rd, wr = mean("dram__bytes_read.sum"), mean("dram__bytes_write.sum")
        e = {"launches_in_capture": len(rs),
             "grid": rs[0][col["Grid Size"]] if "Grid Size" in col else None,
             "dram_bytes_read": rd, "dram_bytes_write": wr,
             "dram_bytes": (rd or 0.0) + (wr or 0.0),
             "gpu_time_ms_under_ncu": mean("gpu__time_duration.sum"),
             "dram_throughput_pct": (mean("dram__bytes_read.sum.pct_of_peak_sustained_elapsed") or 0.0)
                                    + (mean("dram__bytes_write.sum.pct_of_peak_sustained_elapsed") or 0.0),
             "sm_throughput_pct": mean("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
             "fp64_pipe_pct": mean("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
             "tensor_fp64_ops_pct": mean("sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed"),
             "issue_active_pct": mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "warps_active_pct": mean("sm__warps_active.avg.pct_of_peak_sustained_active"),
             "registers_per_thread": mean("launch__registers_per_thread")}
        m = re.search(r"rankb_(?:async_)?kernel<\w+, *(?:\(int\))?(\d+)", name)
        if m:
            e["updates_per_launch"] = int(m.group(1))
        out["kernels"][name] = e
    json.dump(out, open(a.out, "w"), indent=1)
    for k, e in out["kernels"].items():
        print(k[:90], e["launches_in_capture"], e["dram_bytes"], e["gpu_time_ms_under_ncu"], e["fp64_pipe_pct"])


if __name__ == "__main__":
    main()
