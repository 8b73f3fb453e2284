"""bf16 parity of the benchmarked path AT THE BENCHMARKED SIZE (BASELINE.json configs[1]: 2 x 16 x 768 x 1152), layer by
layer and teacher-forced, against the oracle (VERDICT r1 "next round" item 1).

The oracle (oracle/deepcam_oracle.py, plain torch) is evaluated ONCE at full size in fp32 on the CUDA device (TF32 off) with
its tape switched on: that yields, for every primitive of the reference graph, the input the reference feeds it and the
upstream gradient the reference sends back.  Every dense convolution, transposed convolution, depthwise convolution and
BatchNorm(+ReLU) of the network is then run through the product path (engine -> backend -> C ABI -> tcgen05 / bandwidth
kernels) on the oracle's bf16-rounded input / gradient, at the layer's real shape, so the tile picker takes exactly the
modes the benchmark runs (wide, B-resident, two-segment convT, dilation 6/12/18 zero fill at 48 x 72, 304 -> 256,
1536 -> 2048, concat slices), and is compared with the same torch operator the oracle calls, evaluated in fp32 on the same
rounded operands.  Bound: 2e-2 relative L2 per tensor (north_star, bf16); the worst layer of every quantity is recorded in
gpurun_out/parity_fullsize_layers.json.

The oracle runs on the GPU here only as the CHECKER (it is torch/cuDNN fp32; the product never sees it)."""
import json
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import deepcam_oracle as O  # noqa: E402

from architecture import deeplab_xception as dx  # noqa: E402
from deepcam_b200 import engine as E  # noqa: E402
from deepcam_b200.backend import CudaBackend  # noqa: E402
from utils import losses  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT = os.path.join(os.path.dirname(__file__), "..", "gpurun_out")
TOL = 2e-2                       # north_star: per-layer activations and gradients within 2e-2 relative (bf16)
N, H, W = 2, 768, 1152


def rel(a, b):
    d = torch.linalg.vector_norm((a.float() - b.float()).flatten(), dtype=torch.float64)
    n = torch.linalg.vector_norm(b.float().flatten(), dtype=torch.float64)
    return float(d / (n + 1e-300))


def nhwc_bf16(t_nchw, c_pad=None):
    t = t_nchw.detach().permute(0, 2, 3, 1).to(torch.bfloat16)
    if c_pad is not None and c_pad != t.shape[3]:
        out = torch.zeros(t.shape[:3] + (c_pad,), dtype=torch.bfloat16, device=t.device)
        out[..., :t.shape[3]] = t
        return out
    return t.contiguous()


def nchw_f32(t_nhwc, c=None):
    t = t_nhwc if c is None else t_nhwc[..., :c]
    return t.permute(0, 3, 1, 2).float()


def r16(t):
    """bf16 storage rounding, back in fp32"""
    return t.detach().to(torch.bfloat16).float()


@pytest.fixture(scope="module")
def world():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = O.init_state_dict(16, 3, 16, seed=333)
    x, label = O.synthetic_batch(N, H, W, seed=333)
    P = {}
    for k, v in sd.items():
        v = v.to(DEV)
        if k in O.param_names(sd):
            v.requires_grad_(True)
        P[k] = v
    tape = O.Tape(True)
    logits = O.forward(P, x.to(DEV), train=True, tape=tape)
    loss = O.fp_loss(logits, label.to(DEV), O.class_weights())
    loss.backward()
    net = dx.DeepLabv3_plus(16, 3, 16, _print=False)
    net.load_state_dict(sd)
    net.precision = "bf16"
    net = net.to(DEV).train()
    torch.cuda.synchronize()
    yield dict(sd=sd, P=P, recs=tape.records, logits=logits.detach(), loss=float(loss), net=net, x=x, label=label)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _following_bn(recs, rec):
    for name, r in recs.items():
        if r["kind"] == "bn" and r["inp"] is rec["out"]:
            return name
    return None


def _engine_for(params):
    be = CudaBackend(torch.bfloat16, torch.device(DEV))
    assert be.use_tc, "the GPU box must run the tcgen05 path"
    grads = E.GradStore(params, torch.device(DEV))
    be.fill_zero_flat(grads.begin_backward())
    return be, grads, E.Engine(be, True, grads)


def _check_layer(world, name, rec):
    """Runs one conv / convT / depthwise (+ its BatchNorm and ReLU) forward and backward through the engine and returns
    {quantity: relative error} against the torch operator on the same bf16-rounded operands."""
    recs, net, P = world["recs"], world["net"], world["P"]
    kind = rec["kind"]
    mod = net.get_submodule(name)
    bn_name = _following_bn(recs, rec)
    bn_mod = net.get_submodule(bn_name) if bn_name else None
    has_relu = bn_name is not None and (bn_name + "+relu") in recs
    params = [mod.weight] + ([mod.bias] if mod.bias is not None else []) + ([bn_mod.weight, bn_mod.bias] if bn_mod else [])
    be, grads, eng = _engine_for(params)
    x16 = rec["inp"].detach().to(torch.bfloat16)                     # the oracle's input to this layer, storage-rounded
    first_layer = name == "xception_features.conv1"
    xa = E.Act(x16.permute(0, 2, 3, 1).contiguous(), needs_grad=not first_layer)
    xa.is_relu = True
    bnspec = dx._bn_spec(bn_mod) if bn_mod else None
    co = rec["out"].shape[1]
    pad_logits = name == "upsample.last_deconv.0"
    # ---- product forward ----
    cat = None
    if kind == "dw":
        y = eng.dw(xa, dx._dw_spec(mod))
    elif pad_logits:
        y = eng.conv(xa, dx._conv_spec(mod), out_c=8, out_dtype=torch.float32)      # as DeconvUpsampler._emit does
    else:
        y = eng.conv(xa, dx._conv_spec(mod), bn=bnspec)
    top = y
    if bn_mod is not None:
        rm0, rv0 = bn_mod.running_mean.clone(), bn_mod.running_var.clone()
        out_slice = None
        if name.startswith("aspp") or name in ("upsample.deconv2.0", "conv2"):
            # these BatchNorms write channel slices of the two concat buffers (DX:451, DX:379)
            tot, off = {"a": (1280, 256 * (int(name[4]) - 1) if name.startswith("aspp") else 0),
                        "u": (304, 0), "c": (304, 256)}[name[0]]
            n_, h_, w_, _ = y.shape
            cat = eng.new_act(n_, h_, w_, tot, torch.bfloat16)
            cat.t.zero_()
            out_slice = cat.slice(off, co)
        top = eng.bn(y, bnspec, relu=has_relu, out=out_slice)
    torch.cuda.synchronize()
    res = {}
    # ---- reference forward on the same rounded operands (fp32, TF32 off) ----
    xr = x16.float().requires_grad_(not first_layer)
    wr = r16(P[name + ".weight"]).requires_grad_(True)
    br = P[name + ".bias"].detach().clone().requires_grad_(True) if (name + ".bias") in P else None
    if kind == "dw":
        d = rec["dil"]
        yr = F.conv2d(F.pad(xr, (d, d, d, d)), wr, None, rec["stride"], 0, d, xr.shape[1])
    elif kind == "convT":
        yr = F.conv_transpose2d(xr, wr, None, 2, 1, 1)
    else:
        yr = F.conv2d(xr, wr, br, rec["stride"], rec["pad"], rec["dil"])
    res["fwd"] = rel(nchw_f32(y.t, co), yr)
    if pad_logits:
        assert float(y.t[..., co:].abs().max()) == 0.0
    if bn_mod is not None:
        yb = nchw_f32(y.t).detach().requires_grad_(True)             # BatchNorm is teacher-forced on OUR stored conv output
        gam = P[bn_name + ".weight"].detach().clone().requires_grad_(True)
        bet = P[bn_name + ".bias"].detach().clone().requires_grad_(True)
        rm, rv = rm0.clone(), rv0.clone()
        ar = F.batch_norm(yb, rm, rv, gam, bet, True, O.BN_MOMENTUM, O.BN_EPS)
        if has_relu:
            ar = F.relu(ar)
        res["bn_fwd"] = rel(nchw_f32(top.t), ar)
        res["bn_running_mean"] = rel(bn_mod.running_mean, rm)
        res["bn_running_var"] = rel(bn_mod.running_var, rv)
        gtop = recs[bn_name + ("+relu" if has_relu else "")]["out"].grad
    else:
        gtop = rec["out"].grad
    # ---- backward: the oracle's upstream gradient, storage-rounded ----
    g16 = gtop.detach().to(torch.bfloat16)
    gn = g16.permute(0, 2, 3, 1)
    if cat is not None:
        cat.grad = torch.zeros_like(cat.t)
        cat.grad[..., top.c_off:top.c_off + co] = gn
    elif pad_logits:
        top.grad = nhwc_bf16(gtop, 8)
    else:
        top.grad = gn.contiguous()
    eng.backward()
    torch.cuda.synchronize()
    if bn_mod is not None:
        ar.backward(g16.float())
        res["bn_dy"] = rel(nchw_f32(y.grad), yb.grad)
        res["bn_dgamma"] = rel(grads.view(bn_mod.weight), gam.grad)
        res["bn_dbeta"] = rel(grads.view(bn_mod.bias), bet.grad)
    dy = nchw_f32(y.grad, co).contiguous()                          # the gradient OUR conv backward consumed
    wanted = [wr] + ([xr] if not first_layer else []) + ([br] if br is not None else [])
    got = torch.autograd.grad(yr, wanted, dy)
    res["wgrad"] = rel(grads.view(mod.weight), got[0])
    if not first_layer:
        res["dgrad"] = rel(nchw_f32(xa.grad), got[1])
    if br is not None:
        res["bgrad"] = rel(grads.view(mod.bias), got[-1])
    return res


def test_every_layer_teacher_forced_at_full_size(world):
    recs = world["recs"]
    todo = [(n, r) for n, r in recs.items() if r["kind"] in ("conv", "convT", "dw") and not n.startswith("global_avg_pool")]
    assert len(todo) == 63 + 63 + 2 + 4 + 4 + 2 + 7          # depthwise + pointwise of the 63 separable units, stem, skips,
                                                            # ASPP, fuse / low-level 1x1, the seven decoder layers
    table, worst = {}, {}
    for name, rec in todo:
        res = _check_layer(world, name, rec)
        table[name] = dict(res, shape_in=list(rec["inp"].shape), shape_out=list(rec["out"].shape), kind=rec["kind"])
        for q, v in res.items():
            if q not in worst or v > worst[q][1]:
                worst[q] = (name, v)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity_fullsize_layers.json"), "w") as fh:
        json.dump(dict(config="2x16x768x1152 bf16, teacher-forced per layer vs the fp32 oracle on the same bf16-rounded operands",
                       tolerance=TOL, layers=len(table), worst=worst, table=table), fh, indent=1)
    bad = [(n, q, v) for n, t in table.items() for q, v in t.items() if isinstance(v, float) and not (v < TOL)]
    assert not bad, bad[:20]


def _rounded_params(world, prefix):
    """Oracle parameter dict for one sub-module: conv weights storage-rounded to bf16 (what the tcgen05 / depthwise kernels
    consume), BatchNorm parameters fp32, fresh running statistics; every parameter a leaf that requires grad."""
    Pr = {}
    for k, v in world["sd"].items():
        if not k.startswith(prefix):
            continue
        v = v.to(DEV)
        if k in O.param_names(world["sd"]):
            v = (r16(v) if v.dim() == 4 else v.detach().clone()).requires_grad_(True)
        else:
            v = v.clone()
        Pr[k] = v
    return Pr


def _module_case(world, mod, prefix, inputs, ref_fn, gout, record_as, tol_out=TOL, tol_grad=5e-2):
    """Runs `mod` (our nn.Module) on NCHW inputs through its public forward and compares output, input gradients and every
    parameter gradient with `ref_fn(Pr, *inputs)` (oracle functions on bf16-rounded weights, fp32)."""
    for m in mod.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
    mod.precision = "bf16"
    mod.zero_grad()
    xs = [r16(t).requires_grad_(True) for t in inputs]
    out = mod(*xs)
    g = r16(gout)
    out.backward(g)
    torch.cuda.synchronize()
    Pr = _rounded_params(world, prefix)
    xr = [r16(t).requires_grad_(True) for t in inputs]
    # the reference arithmetic (fp32) on the product's STORAGE format: every tensor materialised between two operators is
    # rounded to bf16 (oracle.set_storage_dtype).  Without it the comparison measures the format, not the kernels: a ReLU
    # mask flips wherever the bf16-rounded pre-activation crosses zero (~0.3 % of the elements), which alone moves a
    # sum-type gradient such as dbeta by sqrt(2 * 0.003) ~ 8e-2 of its norm (measured: 5.7e-2 median in the middle flow).
    old_storage = O.set_storage_dtype(torch.bfloat16)
    try:
        ref = ref_fn(Pr, *xr)
    finally:
        O.set_storage_dtype(old_storage)
    ref.backward(g)
    res = dict(out=rel(out, ref))
    for i, (a, b) in enumerate(zip(xs, xr)):
        res["dx%d" % i] = rel(a.grad, b.grad)
    gerr = {}
    for k, p in mod.named_parameters():
        gerr[k] = rel(p.grad, Pr[prefix + k].grad)
    res["worst_param_grad"] = max(gerr.items(), key=lambda kv: kv[1])
    res["median_param_grad"] = sorted(gerr.values())[len(gerr) // 2]
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity_fullsize_module_%s.json" % record_as), "w") as fh:
        json.dump(dict(res, param_grads=gerr), fh, indent=1)
    assert res["out"] < tol_out, res
    assert all(v < tol_grad for k, v in res.items() if k.startswith("dx")), res
    assert res["worst_param_grad"][1] < tol_grad, res
    return res


@pytest.mark.parametrize("blk", ["block1", "block2", "block3", "block5", "block12", "block19", "block20"])
def test_block_modules_at_real_width(world, blk, monkeypatch):
    """Block (DX:69-122) as a module at its real width and spatial size: three (two) separable units + BatchNorms + the
    residual / strided-skip fork, forward and backward, against oracle._block on the same rounded input and weights.  Within a
    block the bf16 storage rounding of each intermediate compounds over <= 7 kernels: output bound 2e-2, gradients 5e-2."""
    monkeypatch.setenv("DEEPCAM_B200_GRAPHS", "0")
    recs = world["recs"]
    prefix = "xception_features.%s." % blk
    rec = recs["xception_features." + blk]
    spec = [s for s in O.xception_blocks(16)[0] if s[0] == blk][0]
    mod = getattr(world["net"].xception_features, blk)
    _module_case(world, mod, prefix, [rec["inp"]], lambda Pr, x: O._block(Pr, prefix[:-1], spec, x, True, O.Tape())[0],
                 rec["out"].grad, blk)


@pytest.mark.parametrize("idx", [1, 2, 3, 4])
def test_aspp_modules_at_real_width(world, idx, monkeypatch):
    """ASPP_module (DX:282-302) 2048 -> 256 at 48 x 72 with dilation 1 / 6 / 12 / 18 (padding by TMA zero fill)."""
    monkeypatch.setenv("DEEPCAM_B200_GRAPHS", "0")
    recs = world["recs"]
    p = "aspp%d" % idx
    rate = O.aspp_rates(16)[idx - 1]
    mod = getattr(world["net"], p)
    _module_case(world, mod, p + ".", [recs[p + ".atrous_convolution"]["inp"]],
                 lambda Pr, x: O.aspp_branch(Pr, p, x, rate, True), recs[p + ".bn+relu"]["out"].grad, p)


def test_deconv_upsampler_module_at_real_width(world, monkeypatch):
    """DeconvUpsampler (DX:347-383): 4 transposed convs (parity classes + the fused 2x2-tap last_deconv), concat-free skip,
    two 3x3 convs, the biased 1x1 - eight tensor-core layers deep, so the bf16 bound on the logits is 3e-2."""
    monkeypatch.setenv("DEEPCAM_B200_GRAPHS", "0")
    recs = world["recs"]
    y = recs["upsample.deconv1.0"]["inp"]
    ll = recs["bn2+relu"]["out"]
    _module_case(world, world["net"].upsample, "upsample.", [y, ll], lambda Pr, a, b: O.decoder(Pr, a, b, True),
                 recs["upsample.last_deconv.0"]["out"].grad, "upsample", tol_out=3e-2, tol_grad=8e-2)


def test_full_size_loss_against_oracle(world, monkeypatch):
    """End to end at full size in bf16 only the scalar loss is comparable (activations of the 130-layer random-init
    net decorrelate by construction); SURVEY §9.3 expects ~2e-4 at this pixel count.  Eager, captured and replayed step."""
    monkeypatch.setenv("DEEPCAM_B200_GRAPHS", "1")
    net = world["net"]
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
    w = O.class_weights()
    xd, ld = world["x"].to(DEV), world["label"].to(DEV)
    seen = []
    for _ in range(3):
        net.zero_grad()
        out = net(xd)
        loss = losses.fp_loss(out, ld, weight=w, fpw_1=w[1], fpw_2=w[2])
        loss.backward()
        seen.append(float(loss))
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity_fullsize_loss.json"), "w") as fh:
        json.dump(dict(oracle_fp32_loss=world["loss"], ours_bf16=seen, logits_rel=rel(out, world["logits"])), fh, indent=1)
    assert all(abs(v - world["loss"]) < 2e-3 for v in seen), (seen, world["loss"])
