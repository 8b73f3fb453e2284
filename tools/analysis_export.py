#!/usr/bin/env python
"""ncu CSV -> the tables the reference's roofline notebooks consume.

The reference's analysis/utils.py turns an Nsight Compute report into a per-kernel frame with the columns
`Name, Metric Name, Invocations, Metric Value` (import_nsight_metric, analysis/utils.py:57-82: group by kernel and metric, average
over invocations) and an overview frame `Name, Time, Invocations, Time Avg` with times in milliseconds (import_nsight_overview,
analysis/utils.py:85-127), and it takes batch size and pass from the FILE NAME (`*.batchsize_<N>.pass_<forward|backward|...>.*`,
parse_filename_nsight, analysis/utils.py:31-42).  This tool writes exactly those two tables from an `ncu --csv` log of THIS
build (e.g. profiles/r02_launches_step_graph.csv, or any `ncu --csv --metrics ...` log with more metrics), with the same column
names and file-name convention, plus the B200 ceilings the roofline plot needs, so the notebooks' pandas code runs on them
unchanged (nv-nsight-cu-cli and the sqlite export of nsys are not needed):

    python tools/analysis_export.py profiles/r02_launches_step_graph.csv --batchsize 2 --pass training --out gpurun_out/analysis
      -> deepcam.batchsize_2.pass_training.metrics.csv   (Name, Metric Name, Invocations, Metric Value)
         deepcam.batchsize_2.pass_training.overview.csv  (Name, Time [ms], Invocations, Time Avg [ms])
         b200_ceilings.json                              (measured HBM GB/s and bf16 TFLOP/s of MEASURED_PEAKS.json)

Blackwell metric names replace the Volta ones of the notebooks: duration = gpu__time_duration.sum (ns), DRAM bytes =
dram__bytes_read.sum + dram__bytes_write.sum, tensor activity = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed."""
import argparse
import csv
import json
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_ncu_csv(path):
    """Rows of an `ncu --csv` log (the header may be preceded by ==PROF== lines)."""
    with open(path, newline="") as fh:
        rows = [r for r in csv.reader(fh) if len(r) > 10]
    hdr = rows[0]
    ix = {k: hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    out = []
    for r in rows[1:]:
        try:
            val = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        out.append((r[ix["Kernel Name"]], r[ix["Metric Name"]], r[ix["Metric Unit"]], val))
    return out


def short_name(kernel):
    """Kernel name without its argument list (what the notebooks group by after demangling)."""
    k = re.sub(r"^void ", "", kernel)
    depth, cut = 0, len(k)
    for i, ch in enumerate(k):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return k[:cut]


def tables(rows):
    agg = {}
    for kernel, metric, unit, val in rows:
        key = (short_name(kernel), metric)
        cnt, tot = agg.get(key, (0, 0.0))
        agg[key] = (cnt + 1, tot + val)
    metrics = [dict(zip(("Name", "Metric Name", "Invocations", "Metric Value"), (k[0], k[1], c, t / c))) for k, (c, t) in agg.items()]
    overview = []
    for (name, metric), (c, t) in agg.items():
        if metric == "gpu__time_duration.sum":                     # ns -> ms, like import_nsight_overview's 1e-6 factor
            overview.append({"Name": name, "Time": t * 1e-6, "Invocations": c, "Time Avg": t * 1e-6 / c})
    overview.sort(key=lambda d: -d["Time"])
    return metrics, overview


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ncu_csv")
    ap.add_argument("--batchsize", type=int, default=2)
    ap.add_argument("--pass", dest="pass_", default="training")
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "analysis"))
    args = ap.parse_args()
    metrics, overview = tables(read_ncu_csv(args.ncu_csv))
    os.makedirs(args.out, exist_ok=True)
    stem = os.path.join(args.out, "deepcam.batchsize_%d.pass_%s" % (args.batchsize, args.pass_))
    for suffix, recs, cols in ((".metrics.csv", metrics, ("Name", "Metric Name", "Invocations", "Metric Value")),
                               (".overview.csv", overview, ("Name", "Time", "Invocations", "Time Avg"))):
        with open(stem + suffix, "w", newline="") as fh:
            w = csv.DictWriter(fh, fieldnames=cols)
            w.writeheader()
            w.writerows(recs)
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    ceil = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")
    if os.path.exists(peaks_path):
        m = json.load(open(peaks_path))
        ceil = dict(hbm_gbs=m["hbm_gbs"], bf16_tflops=m["bf16_tflops"], bf16_tflops_sustained=m.get("bf16_tflops_sustained"),
                    source="MEASURED_PEAKS.json")
    with open(os.path.join(args.out, "b200_ceilings.json"), "w") as fh:
        json.dump(ceil, fh, indent=1)
    print("%d (kernel, metric) rows, %d kernels -> %s{.metrics,.overview}.csv" % (len(metrics), len(overview), stem))


if __name__ == "__main__":
    main()
