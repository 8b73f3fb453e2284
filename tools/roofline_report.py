#!/usr/bin/env python
"""Per-kernel-class roofline table from bench.py's instrumented pass (gpurun_out/bench_kernel_classes.json):
achieved TFLOP/s or GB/s of every kernel class and of the heaviest layer shapes against the measured peaks
(MEASURED_PEAKS.json when present, else the fallback stated in B200_PROFILING.md), as markdown.
  python tools/roofline_report.py profiles/r01_s4_kernel_classes.json > profiles/r01_s4_roofline_table.md
Times are CUDA-event times around single launches in eager mode (they include ~3-5 us of event/launch overhead per launch,
which matters for the 10-20 us kernels; tools/kbench.py gives back-to-back in-graph times for those)."""
import json
import sys

TENSOR = ("conv_gemm_tc", "conv_wgrad_tc", "conv_gemm_simt", "conv_wgrad_simt")


def main(path):
    d = json.load(open(path))
    pk = d["peaks"]
    print("# Kernel-class roofline (%s)\n" % path)
    print("Step (timed, CUDA graphs): %.2f ms.  Peaks (%s): %.0f TFLOP/s bf16 sustained, %.0f GB/s HBM copy.\n"
          % (d["ms_per_step_timed"], pk["source"], pk["tf_sustained"], pk["hbm"]))
    print("| kernel class | launches/step | ms/step (eager events) | bound | achieved | % of peak |")
    print("|---|---|---|---|---|---|")
    for k, v in d["classes"].items():
        if k in TENSOR:
            print("| %s | %d | %.3f | tensor | %.0f TFLOP/s | %.0f %% |" % (k, v["launches_per_step"], v["ms_per_step"], v["tflops"],
                                                                       100.0 * v["tflops"] / pk["tf_sustained"]))
        else:
            print("| %s | %d | %.3f | HBM | %.0f GB/s | %.0f %% |" % (k, v["launches_per_step"], v["ms_per_step"], v["gbs"],
                                                                  100.0 * v["gbs"] / pk["hbm"]))
    print("\n## Heaviest (kernel, layer shape) pairs\n")
    print("| kernel + shape | launches/step | us/launch | ms/step | TFLOP/s | GB/s | % of its roofline |")
    print("|---|---|---|---|---|---|---|")
    rows = sorted(d["conv_shapes"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:45]
    for k, v in rows:
        name = k.split(" ")[0]
        tens = name in TENSOR
        # a GEMM is held against whichever roofline is lower for its arithmetic intensity
        t_frac = v["tflops"] / pk["tf_sustained"]
        b_frac = v["gbs"] / pk["hbm"]
        frac = max(t_frac, b_frac) if tens else b_frac
        print("| %s | %d | %.1f | %.3f | %.0f | %.0f | %.0f %% |" % (k, v["launches_per_step"], v["us_per_launch"], v["ms_per_step"],
                                                                 v["tflops"], v["gbs"], 100.0 * frac))


if __name__ == "__main__":
    main(sys.argv[1])
