#!/usr/bin/env python
"""One profiled training step for Nsight Compute (three-phase windows like the reference's profile_hdf5_ddp.py,
PR:77-94, but through cudaProfilerStart/Stop of torch instead of pycuda).

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --phase all
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_gemm_tc -c 3 \
      -o gpurun_out/prof_tc python tools/profile_step.py --phase forward
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phase", default="all", choices=["all", "forward", "backward", "optimizer"])
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--height", type=int, default=768)
    ap.add_argument("--width", type=int, default=1152)
    args = ap.parse_args()
    import torch
    from architecture import deeplab_xception as dx
    from utils import losses
    dev = torch.device("cuda:0")
    torch.manual_seed(333)
    net = dx.DeepLabv3_plus(16, 3, 16, _print=False).to(dev).train()
    from deepcam_b200.optim import FusedAdam
    opt = FusedAdam(net.parameters(), lr=1e-3, eps=1e-8, weight_decay=1e-6)      # as in bench.py
    cw = [1.001729912096556, 2.6146112239752224, 1.7164197479589602]
    x = torch.rand(args.batch, 16, args.height, args.width, device=dev)
    label = (torch.rand(args.batch, args.height, args.width, device=dev) > 0.986).long()

    class Window:
        def __init__(self, name):
            self.on = args.phase in ("all", name)

        def __enter__(self):
            if self.on:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()

        def __exit__(self, *a):
            if self.on:
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()

    def step(profile):
        class Null:
            def __enter__(self): pass
            def __exit__(self, *a): pass
        W = Window if profile else (lambda name: Null())
        with W("forward"):
            out = net.forward(x)
            loss = losses.fp_loss(out, label, weight=cw, fpw_1=cw[1], fpw_2=cw[2])
        with W("backward"):
            opt.zero_grad()
            loss.backward()
        with W("optimizer"):
            opt.step()
        return loss

    for _ in range(args.warmup):
        step(False)
    loss = step(True)
    torch.cuda.synchronize()
    print("profiled step done, loss", float(loss))


if __name__ == "__main__":
    main()
