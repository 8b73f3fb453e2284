#!/usr/bin/env python
"""Micro-benchmark of single kernels on the network's real layer shapes.

Each case is captured into a CUDA graph of `--reps` back-to-back launches (so host launch cost is excluded) and
timed with CUDA events; prints microseconds per launch and the achieved algorithmic GB/s or TFLOP/s.
  python tools/kbench.py                 # all cases
  python tools/kbench.py --only dw,bn    # substring filter
  ncu ... python tools/kbench.py --only bn_stats --reps 1 --no-graph      # for profiling one kernel
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    from deepcam_b200.backend import BnSpec, ConvSpec, CudaBackend, DwSpec
    dev = torch.device("cuda:0")
    be = CudaBackend(torch.bfloat16, dev)
    bf = torch.bfloat16
    cases = []

    def rnd(*shape):
        return torch.randn(*shape, device=dev).to(bf)

    def add(name, fn, nbytes=0.0, flops=0.0):
        if args.only and not any(s in name for s in args.only.split(",")):
            return
        cases.append((name, fn, nbytes, flops))

    for (n, h, w, c) in [(2, 48, 72, 728), (2, 192, 288, 256), (2, 384, 576, 128)]:
        tag = "%dx%dx%dx%d" % (n, h, w, c)
        y, res, out, dout, dy, dres = (rnd(n, h, w, c) for _ in range(6))
        bn = torch.nn.BatchNorm2d(c).to(dev)
        spec = BnSpec("bn", bn)
        nb = y.numel() * 2.0
        state = {}

        def f_fwd(y=y, res=res, out=out, spec=spec, state=state):
            state["sums"] = be.bn_fwd(y, spec, True, None, out, training=True)
        add("bn_fwd(stats+apply) " + tag, f_fwd, 3 * nb)

        def f_fwd_res(y=y, res=res, out=out, spec=spec):
            be.bn_fwd(y, spec, True, res, out, training=True)
        add("bn_fwd_res " + tag, f_fwd_res, 4 * nb)
        sums0 = be.bn_fwd(y, spec, True, None, out, training=True)

        def f_bwd(dout=dout, out=out, y=y, spec=spec, sums0=sums0, dy=dy):
            be.bn_bwd(dout, out, y, spec, sums0, True, dy, None, False, None, None)
        add("bn_bwd(reduce+apply) " + tag, f_bwd, 7 * nb)

        wdw = torch.nn.Parameter(torch.randn(c, 1, 3, 3, device=dev) * 0.3)
        dws = DwSpec("dw", wdw, 1, 1)
        add("dw_fwd " + tag, lambda y=y, out=out, dws=dws: be.dw_fwd(y, dws, out), 2 * nb, 18.0 * y.numel())
        add("dw_bwd_data " + tag, lambda dout=dout, dy=dy, dws=dws: be.dw_bwd_data(dout, dws, dy, False), 2 * nb, 18.0 * y.numel())
        from deepcam_b200 import ops as _ops
        ws_ready = torch.zeros(_ops.bn_ws_elems(c), dtype=torch.float64, device=dev)
        y64 = y.double().reshape(-1, c)
        ws_ready[:c] = y64.sum(0)
        ws_ready[c:2 * c] = (y64 * y64).sum(0)
        del y64
        add("dw_fwd_bn " + tag, lambda y=y, spec=spec, dws=dws, res=res, out=out, ws=ws_ready: be.bn_dw_fwd(y, spec, True, dws, res, out, ws),
            3 * nb, 21.0 * y.numel())
        add("bn_apply(sums ready) " + tag, lambda y=y, spec=spec, out=out, ws=ws_ready: be.bn_fwd(y, spec, True, None, out, True, ready_sums=ws), 2 * nb)
        rws_holder = {}

        def f_bnred(dout=dout, dy=dy, dws=dws, y=y, sums0=sums0, h=rws_holder):
            h["rws"] = be.dw_bwd_data_bnred(dout, dws, dy, False, y, None, sums0, True, force=True)
        add("dw_bwd_data_bnred " + tag, f_bnred, 3 * nb, 22.0 * y.numel())
        f_bnred()

        def f_bwd_reduced(dout=dout, y=y, spec=spec, sums0=sums0, dy=dres, h=rws_holder):
            be.bn_bwd_reduced(dout, y, spec, sums0, h["rws"], dy, None, False, None, None)
        add("bn_bwd_apply_reduced " + tag, f_bwd_reduced, 3 * nb)
        wg = torch.zeros(c, 1, 3, 3, device=dev)
        add("dw_bwd_weight " + tag, lambda y=y, dout=dout, dws=dws, wg=wg: be.dw_bwd_weight(y, dout, dws, wg), 2 * nb, 18.0 * y.numel())

    for (n, h, w, ci, co, k, dil) in [(2, 48, 72, 728, 728, 1, 1), (2, 192, 288, 256, 256, 1, 1), (2, 48, 72, 2048, 256, 3, 6),
                                      (2, 192, 288, 256, 256, 3, 1), (2, 48, 72, 1536, 2048, 1, 1), (2, 384, 576, 128, 128, 1, 1)]:
        tag = "%dx%dx%d %d->%d k%d" % (n, h, w, ci, co, k)
        x = rnd(n, h, w, ci)
        wt = torch.nn.Parameter(torch.randn(co, ci, k, k, device=dev) * 0.05)
        cs = ConvSpec("c", wt, None, 1, dil * (k // 2), dil)
        out = torch.empty(n, h, w, co, device=dev, dtype=bf)
        dyc = rnd(n, h, w, co)
        dx = torch.empty(n, h, w, ci, device=dev, dtype=bf)
        gw = torch.zeros(co, ci, k, k, device=dev)
        fl = 2.0 * n * h * w * ci * co * k * k
        nb = (x.numel() + out.numel()) * 2.0 + wt.numel() * 2.0
        add("conv_fprop " + tag, lambda x=x, cs=cs, out=out: be.conv_fwd(x, cs, out), nb, fl)
        add("conv_dgrad " + tag, lambda dyc=dyc, cs=cs, dx=dx: be.conv_bwd_data(dyc, cs, dx, False), nb, fl)
        add("conv_wgrad " + tag, lambda x=x, dyc=dyc, cs=cs, gw=gw: be.conv_bwd_weight(x, dyc, cs, gw), nb, fl)

    # fixed cost of the GEMM kernel at the middle-flow tile count: the same 54 x 2 tiles with 1, 6 and 12 k blocks
    for ci in (64, 384, 728):
        x = rnd(2, 48, 72, ci)
        wt = torch.nn.Parameter(torch.randn(728, ci, 1, 1, device=dev) * 0.05)
        cs = ConvSpec("c", wt, None, 1, 0, 1)
        out = torch.empty(2, 48, 72, 728, device=dev, dtype=bf)
        add("kscan_fprop 2x48x72 %d->728 k1" % ci, lambda x=x, cs=cs, out=out: be.conv_fwd(x, cs, out),
            (x.numel() + out.numel() + wt.numel()) * 2.0, 2.0 * 6912 * ci * 728)

    # small-channel layers of the entry flow / last deconv (K or N far below a tensor-core tile)
    for (n, h, w, ci, co, k, stride, transposed, name) in [(2, 768, 1152, 16, 32, 3, 2, False, "conv1"), (2, 384, 576, 32, 64, 3, 1, False, "conv2"),
                                                           (2, 384, 576, 256, 3, 3, 2, True, "last_deconv")]:
        tag = "%s %dx%dx%d %d->%d k%d s%d" % (name, n, h, w, ci, co, k, stride)
        x = rnd(n, h, w, ci)
        if transposed:
            wt = torch.nn.Parameter(torch.randn(ci, co, k, k, device=dev) * 0.05)
            cs = ConvSpec("c", wt, None, stride, 1, 1, True)
            ho, wo = h * 2, w * 2
            out = torch.empty(n, ho, wo, 8, device=dev, dtype=torch.float32)
        else:
            wt = torch.nn.Parameter(torch.randn(co, ci, k, k, device=dev) * 0.05)
            cs = ConvSpec("c", wt, None, stride, 1, 1)
            ho, wo = cs.out_hw(h, w)
            out = torch.empty(n, ho, wo, co, device=dev, dtype=bf)
        fl = 2.0 * n * (h if transposed else ho) * (w if transposed else wo) * ci * co * k * k
        nb = x.numel() * 2.0 + out.numel() * out.element_size() + wt.numel() * 2.0
        add("small_fprop " + tag, lambda x=x, cs=cs, out=out: be.conv_fwd(x, cs, out), nb, fl)
        dys = rnd(n, ho, wo, 8 if transposed else co)
        gws = torch.zeros_like(wt)
        add("small_wgrad " + tag, lambda x=x, dys=dys, cs=cs, gws=gws: be.conv_bwd_weight(x, dys, cs, gws), nb, fl)

    if not args.only or "pack" in args.only:
        # all weight packs of the real network in one launch (what the forward graph starts with)
        from architecture import deeplab_xception as dx
        from deepcam_b200 import ops
        from deepcam_b200.backend import pack_jobs_of
        net = dx.DeepLabv3_plus(16, 3, 16, _print=False).to(dev).train()
        xs = torch.rand(2, 16, 768, 1152, device=dev)      # full size: every layer takes the tcgen05 (NTK pack) path
        net(xs).sum().backward()                       # eager call: creates every kernel-layout weight copy
        jobs = pack_jobs_of(net)
        table = ops.build_pack_table(jobs, dev)
        src_b = sum(j[0].numel() * 4 for j in jobs)
        dst_b = sum(j[1].numel() * j[1].element_size() for j in jobs)
        add("pack_weights_multi (%d jobs)" % len(jobs), lambda table=table: ops.pack_weights_multi(*table), float(src_b + dst_b))

    results = {}
    for name, fn, nbytes, flops in cases:
        fn()
        torch.cuda.synchronize()
        if args.no_graph:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = 1000.0 * e0.elapsed_time(e1) / args.reps
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(args.reps):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = 1000.0 * e0.elapsed_time(e1) / args.reps
        results[name] = dict(us=us, gbs=nbytes / us / 1e3 if nbytes else None, tflops=flops / us / 1e6 if flops else None)
        print("%-44s %8.1f us  %8.0f GB/s  %8.1f TF/s" % (name, us, nbytes / us / 1e3, flops / us / 1e6), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
