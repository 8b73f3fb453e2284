#!/usr/bin/env python
"""Per-tile timeline of the halo mode of the persistent tcgen05 GEMM (CTA 0, first 32 tiles), from the -DDC_TC_TRACE build
(tools/tc_trace.py --build).  For each tile: when the halo load was issued, when the producers saw it / finished, when the MMA
lane owned / committed the accumulator, when the epilogue saw / released / finished it - microseconds since the first stamp.
  python tools/tc_trace.py --build; (GPU box) python tools/halo_trace.py --out gpurun_out/halo_trace.json"""
import argparse
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))
SLOTS = ["halo_issued", "producer_saw_halo", "producer_done", "mma_owns_acc", "mma_committed", "epi_saw_acc", "epi_released_acc", "epi_done"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import numpy as np
    import torch
    from deepcam_b200 import _lib, build as B
    B.LIB_PATH = os.path.join(REPO, "tools", "_trace", "libdeepcam_b200_trace.so")
    B.build = lambda *a, **k: B.LIB_PATH                    # load the instrumented library as it is
    lib = _lib.load()
    lib.dc_halo_trace_read.restype = ctypes.c_int
    lib.dc_halo_trace_read.argtypes = [ctypes.c_void_p]
    from deepcam_b200.backend import ConvSpec, CudaBackend
    dev = torch.device("cuda:0")
    be = CudaBackend(dtype=torch.bfloat16, device=dev, use_tc=True)
    out = []
    for name, (n, h, w, ci, co, stride) in {"conv1 16->32 s2": (2, 768, 1152, 16, 32, 2), "conv2 32->64 s1": (2, 384, 576, 32, 64, 1)}.items():
        x = torch.randn(n, h, w, ci, device=dev).bfloat16()
        wt = torch.nn.Parameter(torch.randn(co, ci, 3, 3, device=dev) * 0.05)
        spec = ConvSpec("c", wt, None, stride, 1, 1)
        ho, wo = spec.out_hw(h, w)
        y = torch.empty(n, ho, wo, co, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            be.conv_fwd(x, spec, y)
        torch.cuda.synchronize()
        buf = np.zeros(32 * 8 + 8, dtype=np.uint64)
        assert lib.dc_halo_trace_read(buf.ctypes.data) == 0
        t = buf[:256].reshape(32, 8).astype(np.int64)
        kbp = buf[256:261].astype(np.int64)
        print(name, "one k block of a producer warp (tile 6), clk: wait_empty %d, copy %d, fence %d, arrive %d" % tuple(int(kbp[k + 1] - kbp[k]) for k in range(4)))
        t0 = t[t > 0].min()
        us = (t - t0) / 1965.0
        rows = [{s: round(float(us[i, j]), 2) for j, s in enumerate(SLOTS)} for i in range(24) if t[i, 7] > 0]
        per_tile = float(us[20, 7] - us[4, 7]) / 16.0 if t[20, 7] > 0 else None
        print(name, "steady-state us per tile:", per_tile)
        for i, r in enumerate(rows[:12]):
            print(i, r)
        out.append(dict(case=name, us_per_tile_steady=per_tile, tiles=rows))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
