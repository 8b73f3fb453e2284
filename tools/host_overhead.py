#!/usr/bin/env python
"""Host-issue time vs device time of one training step (is the step launch-bound?).

For each phase (forward+loss, backward, optimizer) prints the wall time the Python thread needs to ISSUE the work
(no synchronisation inside) and the device time between CUDA events around the phase.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
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
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, eps=1e-8, weight_decay=1e-6)
    cw = [1.001729912096556, 2.6146112239752224, 1.7164197479589602]
    x = torch.rand(args.batch, 16, args.height, args.width, device=dev)
    label = (torch.rand(args.batch, args.height, args.width, device=dev) > 0.986).long()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step(rec):
        t0 = time.perf_counter(); e0 = ev()
        out = net.forward(x)
        loss = losses.fp_loss(out, label, weight=cw, fpw_1=cw[1], fpw_2=cw[2])
        t1 = time.perf_counter(); e1 = ev()
        opt.zero_grad()
        loss.backward()
        t2 = time.perf_counter(); e2 = ev()
        opt.step()
        t3 = time.perf_counter(); e3 = ev()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if rec is not None:
            rec.append(dict(host_fwd_ms=1e3 * (t1 - t0), host_bwd_ms=1e3 * (t2 - t1), host_opt_ms=1e3 * (t3 - t2),
                            dev_fwd_ms=e0.elapsed_time(e1), dev_bwd_ms=e1.elapsed_time(e2), dev_opt_ms=e2.elapsed_time(e3),
                            wall_ms=1e3 * (t4 - t0)))

    for _ in range(3):
        step(None)
    rec = []
    for _ in range(args.steps):
        step(rec)
    avg = {k: sum(r[k] for r in rec) / len(rec) for k in rec[0]}
    print(json.dumps(avg))


if __name__ == "__main__":
    main()
