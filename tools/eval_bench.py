#!/usr/bin/env python
"""Eval-path throughput (BASELINE.json configs[3]): eval() + no_grad forward, argmax and IoU counters over a synthetic
validation set, as the reference's validation loop does (TR:428-492: batch 1, `torch.max(outputs, 1)[1]`,
`compute_score`), on one GPU or data-parallel over the ranks of a torchrun launch (each rank scores its shard, three
scalars are all-reduced at the end like TR:490-492).  Prints one JSON line on rank 0.
  python tools/eval_bench.py [--batch 1] [--samples 64]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/eval_bench.py
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mlperf-deepcam_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--samples", type=int, default=64, help="validation samples per rank")
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from architecture import deeplab_xception as dx
    from utils import utils as dcutils
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(333)
    net = dx.DeepLabv3_plus(n_input=16, n_classes=3, os=16, _print=False).to(dev).eval()
    g = torch.Generator().manual_seed(1000 + rank)
    x = torch.rand((args.batch, 16, 768, 1152), generator=g).to(dev)
    label = (torch.rand((args.batch, 768, 1152), generator=g) > 0.986).long().to(dev)
    iters = max(1, args.samples // args.batch)
    score_sum = torch.zeros((), device=dev)

    def one():
        with torch.no_grad():
            out = net.forward(x)
            return dcutils.argmax_score(out, label, 3)          # argmax + tp/fp/fn + IoU in one pass over the logits

    for _ in range(args.warmup):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        score_sum += one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    stats = torch.tensor([ms, float(score_sum), float(iters)], device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)             # the three scalars of TR:490-492
        ms = float(mx[0])
    if rank == 0:
        total = world * iters * args.batch
        print(json.dumps(dict(metric="eval_samples_per_s_768x1152x16", value=total / (ms / 1000.0), unit="samples/s", n_gpus=world,
                              batch=args.batch, samples=total, ms_per_batch=ms / iters, mean_iou=float(stats[1]) / float(stats[2]),
                              precision=getattr(net, "precision", None) or os.environ.get("DEEPCAM_B200_PRECISION", "bf16"),
                              data="synthetic", path="eval() + no_grad forward (CUDA-graph plan) + fused argmax/IoU kernel")))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
