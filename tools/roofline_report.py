#!/usr/bin/env python
"""Per-kernel-class roofline table from bench.py's instrumented pass (gpurun_out/bench_kernel_classes.json), as markdown:
achieved TFLOP/s or GB/s of every kernel class and of the heaviest (kernel, layer shape) pairs against the peaks MEASURED on
this pool's B200s (MEASURED_PEAKS.json at the repo root; the file's own `peaks` entry is used only when that is absent).
  python tools/roofline_report.py profiles/r02_kernel_classes.json > profiles/r02_roofline_table.md
Two clocks per row: `in graph` = the launch replayed 10x back to back inside a CUDA graph (how it executes in the timed
region; the headline), `eager` = one CUDA event pair around every launch of an eager step (adds ~5-7 us of event + launch
overhead per launch).  A GEMM is held against the lower of its two rooflines at its arithmetic intensity."""
import json
import os
import sys

TENSOR = ("conv_gemm_tc", "conv_wgrad_tc", "conv_gemm_simt", "conv_wgrad_simt")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks(d):
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        m = json.load(open(p))
        return dict(tf=m.get("bf16_tflops_sustained", m["bf16_tflops"]), hbm=m["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    pk = d["peaks"]
    return dict(tf=pk["tf_sustained"], hbm=pk["hbm"], source=pk["source"])


def main(path):
    d = json.load(open(path))
    pk = peaks(d)
    print("# Kernel-class roofline (%s)\n" % path)
    print("Step (timed, CUDA graphs): %.2f ms.  Peaks, %s: %.0f TFLOP/s bf16 sustained, %.0f GB/s HBM copy.\n"
          % (d["ms_per_step_timed"], pk["source"], pk["tf"], pk["hbm"]))
    print("| kernel class | launches/step | ms/step in graph | ms/step eager events | bound | achieved (in graph) | % of measured peak |")
    print("|---|---|---|---|---|---|---|")
    for k, v in d["classes"].items():
        g = v.get("ms_per_step_in_graph")
        ms = g if g else v["ms_per_step"]
        if k in TENSOR:
            ach = v.get("flops_per_step", 0.0) / (ms / 1e3) / 1e12 if g else v["tflops"]
            print("| %s | %d | %s | %.3f | tensor | %.0f TFLOP/s | %.0f %% |" % (k, v["launches_per_step"], "%.3f" % g if g else "-",
                                                                              v["ms_per_step"], ach, 100.0 * ach / pk["tf"]))
        else:
            ach = v.get("bytes_per_step", 0.0) / (ms / 1e3) / 1e9 if g else v["gbs"]
            print("| %s | %d | %s | %.3f | HBM | %.0f GB/s | %.0f %% |" % (k, v["launches_per_step"], "%.3f" % g if g else "-",
                                                                       v["ms_per_step"], ach, 100.0 * ach / pk["hbm"]))
    print("\n## Heaviest (kernel, layer shape) pairs, by in-graph time per step\n")
    print("| kernel + shape | launches/step | us/launch in graph | us/launch eager | ms/step | TFLOP/s | GB/s | % of its roofline |")
    print("|---|---|---|---|---|---|---|---|")

    def t_us(v):
        return v.get("us_per_launch_in_graph") or v["us_per_launch"]

    rows = sorted(d["conv_shapes"].items(), key=lambda kv: -t_us(kv[1]) * kv[1]["launches_per_step"])[:50]
    for k, v in rows:
        name = k.split(" ")[0]
        us = t_us(v)
        tf = v.get("flops_per_launch", 0.0) / us / 1e6 if "flops_per_launch" in v else v["tflops"]
        gb = v.get("bytes_per_launch", 0.0) / us / 1e3 if "bytes_per_launch" in v else v["gbs"]
        frac = max(tf / pk["tf"], gb / pk["hbm"]) if name in TENSOR else gb / pk["hbm"]
        print("| %s | %d | %.1f | %.1f | %.3f | %.0f | %.0f | %.0f %% |" % (k, v["launches_per_step"], us, v["us_per_launch"],
                                                                        us * v["launches_per_step"] / 1e3, tf, gb, 100.0 * frac))


if __name__ == "__main__":
    main(sys.argv[1])
