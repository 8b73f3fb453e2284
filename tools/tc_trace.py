#!/usr/bin/env python
"""Phase timeline of the persistent tcgen05 GEMM kernel (conv_gemm_tc2_kernel) per CTA, from SM-clock timestamps that a
-DDC_TC_TRACE build of csrc/tc_conv.cu records (the product library compiles them away).

  python tools/tc_trace.py --build                     # here (no GPU): tools/_trace/libdeepcam_b200_trace.so
  python tools/tc_trace.py --out gpurun_out/tc_trace.json    # on the GPU box

Shapes: the middle-flow 728 -> 728 pointwise GEMM at 2x48x72 (with and without the BatchNorm-statistics epilogue) launched
alone and as a back-to-back chain (programmatic dependent launch), plus the decoder 3x3."""
import argparse
import ctypes
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))
TRACE_DIR = os.path.join(REPO, "tools", "_trace")
TRACE_LIB = os.path.join(TRACE_DIR, "libdeepcam_b200_trace.so")
SLOTS = ["entry", "after_pdl_wait", "first_stage_landed", "last_mma_issued", "accumulator_complete", "epilogue_done", "exit",
         "globaltimer", "chunk0_staged", "chunk0_store_issued", "chunk0_stats_done", "epilogue_done_group_b",
         "accumulator_complete_group_b"]


def build():
    from deepcam_b200 import build as B
    B.build()
    os.makedirs(TRACE_DIR, exist_ok=True)
    obj = os.path.join(TRACE_DIR, "tc_conv_trace.o")
    subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-DDC_TC_TRACE", "-c", os.path.join(B.CSRC, "tc_conv.cu"), "-o", obj], check=True)
    objs = [os.path.join(B.LIB_DIR, "obj", s.replace(".cu", ".o")) for s in B.SOURCES if s != "tc_conv.cu"] + [obj]
    subprocess.run([B._nvcc(), "-shared", "-o", TRACE_LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"],
                   check=True)
    print(TRACE_LIB)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--out", default="")
    ap.add_argument("--chain", type=int, default=8)
    args = ap.parse_args()
    if args.build:
        return build()
    import numpy as np
    import torch
    from deepcam_b200 import _lib, build as B
    B.LIB_PATH = TRACE_LIB                                  # load the instrumented library instead of the product one
    lib = _lib.load()
    lib.dc_tc_trace_read.restype = ctypes.c_int
    lib.dc_tc_trace_read.argtypes = [ctypes.c_void_p]
    from deepcam_b200.backend import ConvSpec, CudaBackend
    dev = torch.device("cuda:0")
    be = CudaBackend(dtype=torch.bfloat16, device=dev, use_tc=True)
    prop = torch.cuda.get_device_properties(0)
    ghz = 1.965            # B200 boost clock; timestamps are SM clocks (nvidia-smi reports the idle clock outside a kernel)

    def read():
        torch.cuda.synchronize()
        buf = np.zeros(148 * 16, dtype=np.uint64)
        assert lib.dc_tc_trace_read(buf.ctypes.data) == 0
        return buf.reshape(148, 16).astype(np.int64)

    def case(name, n, h, w, ci, co, k, pad, want_sums, chain):
        x = torch.randn(n, h, w, ci, device=dev).bfloat16()
        wt = torch.nn.Parameter(torch.randn(co, ci, k, k, device=dev) * 0.05)
        spec = ConvSpec("c", wt, None, 1, pad, 1)
        outs = [torch.empty(n, h, w, co, device=dev, dtype=torch.bfloat16) for _ in range(2)]
        for _ in range(3):
            be.conv_fwd(x, spec, outs[0], want_bn_sums=want_sums)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for i in range(chain):
            be.conv_fwd(x, spec, outs[i & 1], want_bn_sums=want_sums)
        ev[1].record()
        t = read()
        ctas = int((t[:, 6] > 0).sum())
        t = t[:ctas]
        rel = (t[:, [0, 1, 2, 3, 4, 8, 9, 10, 5, 11, 6]] - t[:, [0]]) / (ghz * 1000.0)          # us since this CTA's entry
        names = ["entry", "after_pdl_wait", "first_stage_landed", "last_mma_issued", "accumulator_complete", "chunk0_staged",
                 "chunk0_store_issued", "chunk0_stats_done", "epilogue_done", "epilogue_done_group_b", "exit"]
        res = dict(case=name, chain=chain, ctas=ctas, us_per_launch_events=ev[0].elapsed_time(ev[1]) * 1000.0 / chain,
                   entry_skew_us=float(t[:, 7].max() - t[:, 7].min()) / 1000.0, sm_ghz_assumed=ghz,
                   median_us_since_entry={nm: round(float(np.median(rel[:, j])), 2) for j, nm in enumerate(names)},
                   max_us_since_entry={nm: round(float(rel[:, j].max()), 2) for j, nm in enumerate(names)})
        print(json.dumps(res))
        return res

    out = []
    for chain in (1, args.chain):
        out.append(case("pw728 2x48x72", 2, 48, 72, 728, 728, 1, 0, False, chain))
        out.append(case("pw728 2x48x72 +bnstats", 2, 48, 72, 728, 728, 1, 0, True, chain))
    out.append(case("dec3x3 2x192x288 256->256 +bnstats", 2, 192, 288, 256, 256, 3, 1, True, 1))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(dict(device=prop.name, slots=SLOTS, cases=out), fh, indent=1)


if __name__ == "__main__":
    main()
