#!/usr/bin/env python
"""Critical-path share of each kernel class inside the captured (CUDA-graph) training step.

Per-launch CUDA events cannot be placed inside a graph, and eager per-launch times include event overhead and hide
the overlap of the weight-gradient branch.  This tool instead re-runs the step with one kernel class NOT launched
(DEEPCAM_B200_ABLATE, see deepcam_b200/ops.py) and reports the step-time difference, plus the split of the intact
step into forward+loss / backward / optimizer measured with CUDA events around the phases.
  python tools/ablate.py [--classes dw_fwd,bn_fwd_onepass,...] [--steps 10]
Each configuration runs in a fresh subprocess (the environment variable is read at import time).
"""
import argparse
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))

CLASSES = ["conv_gemm_", "conv_wgrad_", "dw_fwd", "dw_bwd_data", "dw_bwd_weight", "bn_fwd_onepass", "bn_bwd_onepass",
           "bn_stats", "bn_apply", "bn_bwd_reduce", "bn_bwd_apply", "pack_weights_multi", "copy_view"]


def child(steps):
    import torch
    from architecture import deeplab_xception as dx
    from utils import losses
    from deepcam_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    torch.manual_seed(333)
    net = dx.DeepLabv3_plus(n_input=16, n_classes=3, os=16, _print=False).to(dev).train()
    opt = FusedAdam(net.parameters(), lr=1e-3, eps=1e-8, weight_decay=1e-6)
    cw = [1.001729912096556, 2.6146112239752224, 1.7164197479589602]
    x = torch.rand(2, 16, 768, 1152, device=dev)
    label = (torch.rand(2, 768, 1152, device=dev) > 0.98).long()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]

    def step(e=None):
        if e: e[0].record()
        out = net.forward(x)
        loss = losses.fp_loss(out, label, weight=cw, fpw_1=cw[1], fpw_2=cw[2])
        opt.zero_grad()
        if e: e[1].record()
        loss.backward()
        if e: e[2].record()
        opt.step()
        if e: e[3].record()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        step(ev[i])
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    f = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
    b = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
    o = sum(e[2].elapsed_time(e[3]) for e in ev) / steps
    print(json.dumps(dict(ms=ms, fwd_loss=f, bwd=b, opt=o)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", default=",".join(CLASSES))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    if args.child:
        child(args.steps)
        return
    res = {}
    for cls in [""] + [c for c in args.classes.split(",") if c]:
        env = dict(os.environ, DEEPCAM_B200_ABLATE=cls)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--steps", str(args.steps)], env=env,
                           capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            d = dict(error=(r.stderr or r.stdout)[-300:])
        res[cls or "(intact)"] = d
        base = res["(intact)"].get("ms")
        if "ms" in d and base:
            print("%-22s step %7.3f ms  (fwd+loss %6.3f  bwd %6.3f  opt %5.3f)   delta %+7.3f ms" %
                  (cls or "(intact)", d["ms"], d["fwd_loss"], d["bwd"], d["opt"], d["ms"] - base), flush=True)
        else:
            print(cls, d, flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
