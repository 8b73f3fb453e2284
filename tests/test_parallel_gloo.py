"""Host logic of deepcam_b200.parallel.DistributedDataParallel on CPU: world_size 2, gloo backend, layer math by the
TEST-ONLY torch interpreter.  Checks SURVEY §2.3 semantics: C1 initial broadcast, C2 averaged gradients with
per-rank BatchNorm statistics (several buckets, launched from backward), C3 per-forward buffer broadcast."""
import os
import socket
import sys
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "mlperf-deepcam_b200")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(seed):
    from architecture import deeplab_xception as dx
    torch.manual_seed(seed)
    m = dx.Block(16, 32, reps=2, stride=2, start_with_relu=False)
    m.precision = "fp32"
    return m


def _worker(rank, world, port, q):
    try:
        for p in (PKG, HERE):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import torch_backend
        torch_backend.install()
        from deepcam_b200.parallel import DistributedDataParallel

        net = _build(100 + rank)                       # different initial weights per rank
        ddp = DistributedDataParallel(net, bucket_cap_mb=0.002)   # ~500 floats per bucket -> several buckets
        ref0 = _build(100)                             # rank 0's weights
        for (k, a), (_, b) in zip(net.state_dict().items(), ref0.state_dict().items()):
            assert torch.equal(a, b), "C1 broadcast failed for " + k
        assert list(ddp.state_dict().keys())[0].startswith("module.")

        g = torch.Generator().manual_seed(7 + rank)
        x = torch.randn(2, 16, 8, 12, generator=g)
        y = ddp(x)
        y.square().mean().backward()
        assert len(ddp._sync.buckets) > 2

        # expected: average over ranks of the single-process gradients (per-rank BN statistics)
        expect = None
        for r in range(world):
            m = _build(100)
            gr = torch.Generator().manual_seed(7 + r)
            xr = torch.randn(2, 16, 8, 12, generator=gr)
            m(xr).square().mean().backward()
            gs = [p.grad.clone() for p in m.parameters()]
            expect = gs if expect is None else [a + b for a, b in zip(expect, gs)]
        expect = [e / world for e in expect]
        for (k, p), e in zip(net.named_parameters(), expect):
            assert torch.allclose(p.grad, e, rtol=1e-5, atol=1e-7), "C2 mismatch " + k

        # C3: running statistics diverge per rank after the step; the next forward re-broadcasts rank 0's
        stats = net.skipbn.running_mean.clone()
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        assert not torch.equal(gathered[0], gathered[1])
        with torch.no_grad():
            ddp(x)
        # after this forward every rank started from rank 0's buffers; check the counter, which is data independent
        cnt = net.skipbn.num_batches_tracked.clone()
        cg = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cg, cnt)
        assert int(cg[0]) == int(cg[1]) == 2
        # buffers still live in the state_dict under their usual names
        assert "skipbn.running_mean" in net.state_dict()

        # buffers re-homed behind the wrapper's back (module.float() / .to() / load_state_dict(assign=True) install
        # fresh tensors in mod._buffers) must not silently drop out of the C3 broadcast
        assert ddp._buffers_aliased()
        net.skipbn._buffers["running_mean"] = net.skipbn.running_mean.clone() + float(rank + 1)     # ranks now disagree
        assert not ddp._buffers_aliased()
        with torch.no_grad():
            ddp(x)                                       # re-flattens, then broadcasts rank 0's buffers
        assert ddp._buffers_aliased()
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, net.skipbn.running_mean.clone())
        # both ranks started this forward from rank 0's (shifted) statistics and then applied their own batch update;
        # without the re-flatten they would differ by the injected offset of 1.0
        assert float((gathered[0] - gathered[1]).abs().max()) < 0.5

        # a trainable parameter that forward never touches must not stall the in-order bucket launch forever:
        # the first backward learns who reports, later backwards expect only those
        net.unused = torch.nn.Parameter(torch.zeros(3))
        net._dc_gradstore = None
        for it in range(2):
            net.zero_grad()
            ddp(x).square().mean().backward()
            assert ddp._sync._expected is not None and len(ddp._sync._expected) == len(list(net.parameters())) - 1
        ddp._sync.begin(net._dc_gradstore)               # pending counts of the next backward skip the silent parameter
        assert sum(ddp._sync.pending) == len(list(net.parameters())) - 1
        net._dc_gradstore.on_ready = None
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


def test_ddp_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", "rank %d failed:\n%s" % (rank, msg)
