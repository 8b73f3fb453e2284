"""CPU oracle for the DeepCAM training step — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this
module, and only as the checker (or as the timed CPU baseline), never as part of the shipped path.

It is a functional restatement (plain torch fp32/fp64 ops on CPU, no nn.Module classes of the reference,
no code copied) of the reference's hot path:
    DX = /root/reference/src/deepCam/architecture/deeplab_xception.py
    LS = /root/reference/src/deepCam/utils/losses.py
    UT = /root/reference/src/deepCam/utils/utils.py
    TR = /root/reference/src/deepCam/train_hdf5_ddp.py
The arithmetic itself lives in PyTorch (third-party, not vendored in the reference; historical pin
nvcr.io/nvidia/pytorch:20.01-py3, docker/Dockerfile.train:22; here torch 2.11.0) — the oracle calls the same
torch.nn.functional operators the reference's nn.Modules dispatch to.

PARITY PINNING: the reference ships no tests or golden vectors.  The oracle is pinned
against outputs of the reference itself, run in the build container: tests/test_oracle_vs_reference.py compares
it with the live reference classes (bit-exact initial state_dict, logits, loss, all 301 parameter gradients,
running statistics, IoU), and tests/golden/*.npz holds vectors generated from the live reference by
tests/golden/make_golden.py, which travel to the GPU box where /root/reference does not exist.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # torch.nn.BatchNorm2d defaults; DX never overrides them
BN_MOMENTUM = 0.1

# class frequencies hard-coded at TR:206 and the exponent TR:204 (loss_weight_pow default -0.125, TR:571)
CLASS_FREQ = (0.986267818390377, 0.0004578708870701058, 0.01327431072255291)


def class_weights(power=-0.125):
    """TR:204-209: class_weights = [f**p for f in frequencies]."""
    return [f ** power for f in CLASS_FREQ]


# --------------------------------------------------------------------------------------------------
# architecture tables (DX:125-186, 398-439)
# --------------------------------------------------------------------------------------------------
def block_slots(inplanes, planes, reps, stride, dilation, start_with_relu, grow_first, is_last):
    """Slot list of Block.rep with the Sequential indices the state_dict uses (DX:80-109).
    Entries: ("relu",) | ("sep", cin, cout, stride, dilation) | ("bn", c)."""
    slots = []
    filters = inplanes
    if grow_first:
        slots += [("relu",), ("sep", inplanes, planes, 1, dilation), ("bn", planes)]
        filters = planes
    for _ in range(reps - 1):
        slots += [("relu",), ("sep", filters, filters, 1, dilation), ("bn", filters)]
    if not grow_first:
        slots += [("relu",), ("sep", inplanes, planes, 1, dilation), ("bn", planes)]
    if not start_with_relu:
        slots = slots[1:]
    if stride != 1:
        slots.append(("sep", planes, planes, 2, 1))
    if stride == 1 and is_last:
        slots.append(("sep", planes, planes, 1, 1))
    return slots


def xception_blocks(os=16):
    """(name, inplanes, planes, reps, stride, dilation, start_with_relu, grow_first, is_last) per Block, DX:132-177."""
    if os == 16:
        b3_stride, mid_rate, exit_rates = 2, 1, (1, 2)
    elif os == 8:
        b3_stride, mid_rate, exit_rates = 1, 2, (2, 4)
    else:
        raise NotImplementedError
    blocks = [("block1", 64, 128, 2, 2, 1, False, True, False),
              ("block2", 128, 256, 2, 2, 1, True, True, False),
              ("block3", 256, 728, 2, b3_stride, 1, True, True, True)]
    for i in range(4, 20):
        blocks.append(("block%d" % i, 728, 728, 3, 1, mid_rate, True, True, False))
    blocks.append(("block20", 728, 1024, 2, 1, exit_rates[0], True, False, True))
    return blocks, exit_rates[1]


def aspp_rates(os=16):
    return {16: [1, 6, 12, 18], 8: [1, 12, 24, 36]}[os]


# --------------------------------------------------------------------------------------------------
# initialisation: reproduces the RNG consumption order of the reference constructors
# --------------------------------------------------------------------------------------------------
def _default_conv_init(w, bias=None):
    """torch.nn.modules.conv._ConvNd.reset_parameters (what nn.Conv2d / nn.ConvTranspose2d do on construction)."""
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    if bias is not None:
        fan_in, _ = torch.nn.init._calculate_fan_in_and_fan_out(w)
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        torch.nn.init.uniform_(bias, -bound, bound)


def _bn_entries(sd, prefix, c):
    sd[prefix + ".weight"] = torch.ones(c)
    sd[prefix + ".bias"] = torch.zeros(c)
    sd[prefix + ".running_mean"] = torch.zeros(c)
    sd[prefix + ".running_var"] = torch.ones(c)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def init_state_dict(n_input=16, n_classes=3, os=16, seed=None):
    """state_dict (532 entries for the DeepCAM configuration) identical to
    `torch.manual_seed(seed); DeepLabv3_plus(n_input, n_classes, os).state_dict()` of the reference."""
    if seed is not None:
        torch.manual_seed(seed)
    sd = OrderedDict()
    xconvs = []          # Xception conv weights in construction (= modules()) order, re-initialised at DX:189

    def conv(name, co, ci, k, record=None, bias=False, transposed=False):
        w = torch.empty((ci, co, k, k) if transposed else (co, ci, k, k))
        b = torch.empty(co) if bias else None
        _default_conv_init(w, b)
        sd[name + ".weight"] = w
        if bias:
            sd[name + ".bias"] = b
        if record is not None:
            record.append(w)
        return w

    def sep(prefix, cin, cout, record):
        conv(prefix + ".conv1", cin, 1, 3, record)        # depthwise: weight [cin, 1, 3, 3]
        conv(prefix + ".pointwise", cout, cin, 1, record)

    X = "xception_features."
    conv(X + "conv1", 32, n_input, 3, xconvs); _bn_entries(sd, X + "bn1", 32)
    conv(X + "conv2", 64, 32, 3, xconvs); _bn_entries(sd, X + "bn2", 64)
    blocks, exit_rate = xception_blocks(os)
    for (name, cin, cout, reps, stride, dil, swr, gf, last) in blocks:
        p = X + name
        if cout != cin or stride != 1:                      # DX:73-75
            conv(p + ".skip", cout, cin, 1, xconvs); _bn_entries(sd, p + ".skipbn", cout)
        for idx, slot in enumerate(block_slots(cin, cout, reps, stride, dil, swr, gf, last)):
            if slot[0] == "sep":
                sep("%s.rep.%d" % (p, idx), slot[1], slot[2], xconvs)
            elif slot[0] == "bn":
                _bn_entries(sd, "%s.rep.%d" % (p, idx), slot[1])
    sep(X + "conv3", 1024, 1536, xconvs); _bn_entries(sd, X + "bn3", 1536)
    sep(X + "conv4", 1536, 1536, xconvs); _bn_entries(sd, X + "bn4", 1536)
    sep(X + "conv5", 1536, 2048, xconvs); _bn_entries(sd, X + "bn5", 2048)
    for w in xconvs:                                        # Xception.__init_weight, DX:244-252
        torch.nn.init.kaiming_normal_(w)
    for i, rate in enumerate(aspp_rates(os)):               # ASPP_module + its own __init_weight, DX:282-312
        p = "aspp%d" % (i + 1)
        w = conv(p + ".atrous_convolution", 256, 2048, 1 if rate == 1 else 3)
        _bn_entries(sd, p + ".bn", 256)
        torch.nn.init.kaiming_normal_(w)
    conv("global_avg_pool.1", 256, 2048, 1); _bn_entries(sd, "global_avg_pool.2", 256)      # DX:425-428
    conv("conv1", 256, 1280, 1); _bn_entries(sd, "bn1", 256)                                 # DX:430-431
    conv("conv2", 48, 128, 1); _bn_entries(sd, "bn2", 48)                                    # DX:434-435
    U = "upsample."                                                                           # DX:347-374 (default init only)
    conv(U + "deconv1.0", 256, 256, 3, transposed=True); _bn_entries(sd, U + "deconv1.1", 256)
    conv(U + "deconv2.0", 256, 256, 3, transposed=True); _bn_entries(sd, U + "deconv2.1", 256)
    conv(U + "conv1.0", 256, 304, 3); _bn_entries(sd, U + "conv1.1", 256)
    conv(U + "conv1.3", 256, 256, 3); _bn_entries(sd, U + "conv1.4", 256)
    conv(U + "conv1.6", 256, 256, 1, bias=True)
    conv(U + "deconv3.0", 256, 256, 3, transposed=True); _bn_entries(sd, U + "deconv3.1", 256)
    conv(U + "last_deconv.0", n_classes, 256, 3, transposed=True)
    return sd


# --------------------------------------------------------------------------------------------------
# forward (functional).  `P` maps state_dict names to tensors; BN running statistics are updated in place.
# --------------------------------------------------------------------------------------------------
class Tape:
    """Optional record of every primitive's (input, output) for teacher-forced per-layer parity."""

    def __init__(self, enabled=False):
        self.enabled = enabled
        self.records = OrderedDict()

    def add(self, name, kind, inp, out, **attrs):
        if self.enabled:
            if out.requires_grad:
                out.retain_grad()
            self.records[name] = dict(kind=kind, inp=inp, out=out, **attrs)


_STORAGE = None      # None: the reference arithmetic (fp32/fp64 everywhere).  torch.bfloat16: see set_storage_dtype


def set_storage_dtype(dtype):
    """TEST AID.  With a dtype set, every tensor the reference would materialise between two operators (conv / depthwise /
    transposed-conv outputs, BatchNorm(+ReLU) outputs, residual sums) is rounded to that dtype and back (straight-through:
    the gradient passes unchanged), while all arithmetic stays in the tensors' own precision.  This models the product's
    bf16 STORAGE format on top of the reference arithmetic, so that a multi-layer comparison isolates kernel errors from the
    format's own effect (e.g. ReLU masks that flip where a bf16-rounded pre-activation crosses zero).  Returns the old value."""
    global _STORAGE
    old, _STORAGE = _STORAGE, dtype
    return old


def _q(x):
    if _STORAGE is None:
        return x
    return x + (x.detach().to(_STORAGE).to(x.dtype) - x.detach())


def _bn(P, name, x, train, tape, relu=False):
    rm, rv = P[name + ".running_mean"], P[name + ".running_var"]
    if train:
        if x.numel() // x.shape[1] <= 1:
            raise ValueError("Expected more than 1 value per channel when training, got input size %s" % (x.shape,))
        P[name + ".num_batches_tracked"] += 1
    y = F.batch_norm(x, rm, rv, P[name + ".weight"], P[name + ".bias"], train, BN_MOMENTUM, BN_EPS)
    tape.add(name, "bn", x, y)
    if relu:
        y = F.relu(y)
        tape.add(name + "+relu", "relu", None, y)
    return _q(y)


def _sep(P, name, x, stride, dil, tape):
    """SeparableConv2d_same.forward, DX:62-66, with fixed_padding DX:45-51 (kernel 3: pad = dilation on each side)."""
    xp = F.pad(x, (dil, dil, dil, dil))
    c = x.shape[1]
    t = F.conv2d(xp, P[name + ".conv1.weight"], None, stride, 0, dil, c)
    tape.add(name + ".conv1", "dw", x, t, stride=stride, dil=dil)
    t = _q(t)
    y = F.conv2d(t, P[name + ".pointwise.weight"])
    tape.add(name + ".pointwise", "conv", t, y, stride=1, pad=0, dil=1)
    return _q(y)


def _block(P, p, spec, inp, train, tape):
    """Block.forward, DX:111-122.  The reference's first rep element is an in-place ReLU (DX:79,84), so the skip
    path and every alias of `inp` see relu(inp).  Returns (block output, block input as seen afterwards)."""
    (_, cin, cout, reps, stride, dil, swr, gf, last) = spec
    slots = block_slots(cin, cout, reps, stride, dil, swr, gf, last)
    if swr:
        inp = F.relu(inp)                      # in-place in the reference: both paths consume the clamped tensor
    x = inp
    for idx, slot in enumerate(slots):
        name = "%s.rep.%d" % (p, idx)
        if slot[0] == "relu":
            if idx > 0:
                x = F.relu(x)
        elif slot[0] == "sep":
            x = _sep(P, name, x, slot[3], slot[4], tape)
        else:
            x = _bn(P, name, x, train, tape)
    if (p + ".skip.weight") in P:
        s = F.conv2d(inp, P[p + ".skip.weight"], None, stride)
        tape.add(p + ".skip", "conv", inp, s, stride=stride, pad=0, dil=1)
        s = _bn(P, p + ".skipbn", _q(s), train, tape)
    else:
        s = inp
    out = _q(x + s) if _STORAGE is not None else x + s
    tape.add(p, "block", inp, out)
    return out, inp


def aspp_branch(P, p, feats, rate, train=True, tape=None):
    """ASPP_module.forward, DX:298-302: conv (1x1, or 3x3 with dilation = padding = rate) -> BN -> ReLU."""
    tape = tape if tape is not None else Tape(False)
    pad = 0 if rate == 1 else rate
    y = F.conv2d(feats, P[p + ".atrous_convolution.weight"], None, 1, pad, rate)
    tape.add(p + ".atrous_convolution", "conv", feats, y, stride=1, pad=pad, dil=rate)
    return _bn(P, p + ".bn", _q(y), train, tape, relu=True)


def decoder(P, y, ll, train=True, tape=None):
    """DeconvUpsampler.forward, DX:376-383: y = fused ASPP features [N,256,h,w], ll = reduced low-level features [N,48,4h,4w]."""
    tape = tape if tape is not None else Tape(False)
    U = "upsample."

    def deconv(name, v):
        o = F.conv_transpose2d(v, P[U + name + ".0.weight"], None, 2, 1, 1)
        tape.add(U + name + ".0", "convT", v, o)
        return o if name == "last_deconv" else _q(o)          # the logits stay fp32 in the product as well

    y = _bn(P, U + "deconv1.1", deconv("deconv1", y), train, tape, relu=True)
    y = _bn(P, U + "deconv2.1", deconv("deconv2", y), train, tape, relu=True)
    y = torch.cat((y, ll), dim=1)
    t = F.conv2d(y, P[U + "conv1.0.weight"], None, 1, 1)
    tape.add(U + "conv1.0", "conv", y, t, stride=1, pad=1, dil=1)
    t = _q(t)
    y = _bn(P, U + "conv1.1", t, train, tape, relu=True)
    t = F.conv2d(y, P[U + "conv1.3.weight"], None, 1, 1)
    tape.add(U + "conv1.3", "conv", y, t, stride=1, pad=1, dil=1)
    t = _q(t)
    y = _bn(P, U + "conv1.4", t, train, tape, relu=True)
    t = F.conv2d(y, P[U + "conv1.6.weight"], P[U + "conv1.6.bias"])
    tape.add(U + "conv1.6", "conv", y, t, stride=1, pad=0, dil=1)
    t = _q(t)
    y = _bn(P, U + "deconv3.1", deconv("deconv3", t), train, tape, relu=True)
    return deconv("last_deconv", y)


def xception(P, x, train=True, os=16, tape=None):
    """Xception.forward, DX:195-242: returns (features [N,2048,H/os,W/os], low_level_feat [N,128,H/4,W/4])."""
    tape = tape or Tape(False)
    X = "xception_features."
    h = F.conv2d(x, P[X + "conv1.weight"], None, 2, 1)
    tape.add(X + "conv1", "conv", x, h, stride=2, pad=1, dil=1)
    h = _bn(P, X + "bn1", _q(h), train, tape, relu=True)
    t = F.conv2d(h, P[X + "conv2.weight"], None, 1, 1)
    tape.add(X + "conv2", "conv", h, t, stride=1, pad=1, dil=1)
    h = _bn(P, X + "bn2", _q(t), train, tape, relu=True)
    blocks, exit_rate = xception_blocks(os)
    low = None
    outs = []
    for spec in blocks:
        h, seen_inp = _block(P, X + spec[0], spec, h, train, tape)
        outs.append(h)
        if spec[0] == "block2":
            low = seen_inp          # low_level_feat aliases block1's output, clamped by block2's in-place ReLU (DX:206)
    for i, name in enumerate(("conv3", "conv4", "conv5")):
        h = _sep(P, X + name, h, 1, exit_rate, tape)
        h = _bn(P, X + "bn%d" % (i + 3), h, train, tape, relu=True)
    return h, low


def forward(P, x, train=True, os=16, tape=None):
    """DeepLabv3_plus.forward, DX:441-465 (Xception.forward DX:195-242, DeconvUpsampler.forward DX:376-383).
    x: [N, n_input, H, W] with H, W multiples of 16.  Returns logits [N, n_classes, H, W]."""
    tape = tape or Tape(False)
    feats, low = xception(P, x, train, os, tape)
    branches = []
    for i, rate in enumerate(aspp_rates(os)):
        branches.append(aspp_branch(P, "aspp%d" % (i + 1), feats, rate, train, tape))
    g = F.adaptive_avg_pool2d(feats, 1)
    tape.add("global_avg_pool.0", "gap", feats, g)
    g2 = F.conv2d(g, P["global_avg_pool.1.weight"])
    tape.add("global_avg_pool.1", "conv", g, g2, stride=1, pad=0, dil=1)
    g2 = _bn(P, "global_avg_pool.2", g2, train, tape, relu=True)
    g2 = F.interpolate(g2, size=feats.shape[2:], mode="bilinear", align_corners=True)
    cat = torch.cat(branches + [g2], dim=1)
    y = F.conv2d(cat, P["conv1.weight"])
    tape.add("conv1", "conv", cat, y, stride=1, pad=0, dil=1)
    y = _bn(P, "bn1", _q(y), train, tape, relu=True)
    ll = F.conv2d(low, P["conv2.weight"])
    tape.add("conv2", "conv", low, ll, stride=1, pad=0, dil=1)
    ll = _bn(P, "bn2", _q(ll), train, tape, relu=True)
    return decoder(P, y, ll, train, tape)


def interpolation_upsampler(P, x, low, input_size, train=True, tape=None):
    """InterpolationUpsampler.forward, DX:326-333 (the decoder variant DeepLabv3_plus leaves commented out at DX:438).
    P holds the module's own state_dict keys ("last_conv.0.weight", ...).  Pinned against the live reference class by
    tests/test_engine_graph_cpu.py::test_interpolation_upsampler_matches_reference."""
    tape = tape if tape is not None else Tape(False)
    size = (int(math.ceil(input_size[-2] / 4)), int(math.ceil(input_size[-1] / 4)))
    y = F.interpolate(x, size=size, mode="bilinear", align_corners=True)
    y = torch.cat((y, low), dim=1)
    t = F.conv2d(y, P["last_conv.0.weight"], None, 1, 1)
    y = _bn(P, "last_conv.1", t, train, tape, relu=True)
    t = F.conv2d(y, P["last_conv.3.weight"], None, 1, 1)
    y = _bn(P, "last_conv.4", t, train, tape, relu=True)
    t = F.conv2d(y, P["last_conv.6.weight"], P["last_conv.6.bias"])
    return F.interpolate(t, size=tuple(input_size[2:]), mode="bilinear", align_corners=True)


# --------------------------------------------------------------------------------------------------
# loss and metric
# --------------------------------------------------------------------------------------------------
def fp_loss(logit, target, weight, fpw_1=0, fpw_2=0):
    """LS:28-52.  The two re-weighting masks (LS:41, LS:46) are identically zero, so fpw_1/fpw_2 are dead and the
    result is mean over N*H*W of w[t] * (-log softmax(logit)[t]) — a plain mean, not a weight-normalised one."""
    w = torch.as_tensor(np.array(weight), dtype=torch.float32).to(logit.dtype).to(logit.device)
    target = target.squeeze(1).long()
    logp = F.log_softmax(logit, dim=1)
    picked = logp.gather(1, target.unsqueeze(1)).squeeze(1)
    losses = -w[target] * picked
    return losses.mean()


def confusion_counts(prediction, gt, num_classes):
    """Integer tp / fp / fn per class, UT:43-50."""
    gt = gt.long()
    eq = prediction == gt
    ne = ~eq
    tp = [int((eq & (gt == j)).sum()) for j in range(num_classes)]
    fp = [int((ne & (prediction == j)).sum()) for j in range(num_classes)]
    fn = [int((ne & (gt == j)).sum()) for j in range(num_classes)]
    return tp, fp, fn


def compute_score(prediction, gt, num_classes):
    """UT:32-60: mean over classes of tp/(tp+fp+fn) in fp32, 1.0 for an empty union."""
    tp, fp, fn = confusion_counts(prediction, gt, num_classes)
    total = torch.tensor(0.0)
    first = True
    for j in range(num_classes):
        union = tp[j] + fp[j] + fn[j]
        iou = torch.tensor(1.0) if union == 0 else torch.tensor(float(tp[j]), dtype=torch.float32) / torch.tensor(
            float(union), dtype=torch.float32)
        total = iou if first else total + iou
        first = False
    return total / float(num_classes)


# --------------------------------------------------------------------------------------------------
# training step (TR:345-371) and synthetic data
# --------------------------------------------------------------------------------------------------
def synthetic_batch(n, h=768, w=1152, c=16, seed=333, label_mode="freq"):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((n, c, h, w), generator=g)
    if label_mode == "freq":
        u = torch.rand((n, h, w), generator=g)
        label = torch.zeros((n, h, w), dtype=torch.long)
        label[u > CLASS_FREQ[0]] = 2
        label[u > CLASS_FREQ[0] + CLASS_FREQ[2]] = 1
    else:
        label = torch.randint(0, 3, (n, h, w), generator=g)
    return x, label


def param_names(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))]


class TrainState:
    """Parameters + buffers + Adam, stepping exactly like the loop body TR:345-371 (fp32, CPU)."""

    def __init__(self, sd, lr=1e-3, eps=1e-8, weight_decay=1e-6, os=16):
        self.P = OrderedDict()
        for k, v in sd.items():
            v = v.detach().clone()
            if k in param_names(sd):
                v.requires_grad_(True)
            self.P[k] = v
        self.params = [self.P[k] for k in param_names(sd)]
        self.opt = torch.optim.Adam(self.params, lr=lr, eps=eps, weight_decay=weight_decay)
        self.os = os
        self.weights = class_weights()

    def step(self, x, label):
        logits = forward(self.P, x, train=True, os=self.os)
        loss = fp_loss(logits, label, self.weights, self.weights[1], self.weights[2])
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach()), logits.detach()
