"""B200-native drop-in for the reference's `architecture.deeplab_xception` (DX = reference
src/deepCam/architecture/deeplab_xception.py).

Same class names, constructor arguments, attribute paths and therefore the same 532-entry state_dict,
initialisation order and RNG consumption as the reference, so `train_hdf5_ddp.py` (TR:195-201, 227, 237, 352,
521), optimizers and checkpoints work unchanged.  The torch.nn layers created here are *parameter holders
only*: their forward() is never called.  Each class describes its computation to deepcam_b200.engine
(`_emit`), which runs hand-written sm_100a kernels through the C ABI (include/deepcam_b200.h) and supplies
the backward pass; there is no cuDNN/cuBLAS path and no CPU fallback.

Reference quirks reproduced on purpose:
  * Block's leading ReLU is in-place in the reference (DX:79, 84), so the skip branch and `low_level_feat`
    (DX:206) see relu(input).  Here block outputs that feed such a block are stored already clamped.
  * BatchNorm over the 1x1 image-pooling map fails for batch size 1 in train mode (DX:425-428).
"""
import math

import torch
import torch.nn as nn

from deepcam_b200 import engine as _engine
from deepcam_b200.engine import BnSpec, ConvSpec, DwSpec

__all__ = ["SeparableConv2d", "fixed_padding", "SeparableConv2d_same", "Block", "Xception", "ASPP_module",
           "InterpolationUpsampler", "DeconvUpsampler", "DeepLabv3_plus", "get_1x_lr_params", "get_10x_lr_params"]


# ---------------------------------------------------------------------------------------------------------
# helpers shared by the classes below
# ---------------------------------------------------------------------------------------------------------
def _spec_cache(mod):
    cache = mod.__dict__.get("_dc_specs")
    if cache is None:
        cache = {}
        mod.__dict__["_dc_specs"] = cache
    return cache


def _conv_spec(conv, name="conv"):
    """ConvSpec for an nn.Conv2d / nn.ConvTranspose2d parameter holder."""
    cache = _spec_cache(conv)
    spec = cache.get("conv")
    if spec is None or spec.weight is not conv.weight or spec.bias is not conv.bias:
        transposed = isinstance(conv, nn.ConvTranspose2d)
        k, s, p, d = conv.kernel_size, conv.stride, conv.padding, conv.dilation
        if k[0] != k[1] or s[0] != s[1] or p[0] != p[1] or d[0] != d[1] or conv.groups != 1:
            raise NotImplementedError("deepcam_b200: only square, ungrouped dense convolutions are supported (%s)" % conv)
        if transposed and not (k[0] == 3 and s[0] == 2 and p[0] == 1 and tuple(conv.output_padding) == (1, 1)):
            raise NotImplementedError("deepcam_b200: ConvTranspose2d must be k3 s2 p1 output_padding 1 (DX:352)")
        spec = ConvSpec(name, conv.weight, conv.bias, s[0], p[0], d[0], transposed)
        cache["conv"] = spec
    return spec


def _dw_spec(conv, name="dw"):
    cache = _spec_cache(conv)
    spec = cache.get("dw")
    if spec is None or spec.weight is not conv.weight:
        if conv.kernel_size != (3, 3) or conv.bias is not None or conv.groups != conv.in_channels or \
                conv.in_channels != conv.out_channels:
            raise NotImplementedError("deepcam_b200: depthwise convolution must be 3x3, bias-free, groups == channels")
        spec = DwSpec(name, conv.weight, conv.stride[0], conv.dilation[0])
        cache["dw"] = spec
    return spec


def _bn_spec(bn, name="bn"):
    if not isinstance(bn, nn.BatchNorm2d):
        raise NotImplementedError("deepcam_b200: normalizer must be torch.nn.BatchNorm2d, got %s" % type(bn).__name__)
    cache = _spec_cache(bn)
    spec = cache.get("bn")
    if spec is None:
        spec = BnSpec(name, bn)
        cache["bn"] = spec
    return spec


class _EngineModule(nn.Module):
    """forward(x) of every class runs its `_emit` through the engine as a single autograd node."""

    def _emit_root(self, eng, x):
        out = self._emit(eng, x)
        return [(out, out.shape[3])]

    def forward(self, x):
        return _engine.run_module(self, [x])[0]


# ---------------------------------------------------------------------------------------------------------
# DX:31-66
# ---------------------------------------------------------------------------------------------------------
class SeparableConv2d(_EngineModule):
    """DX:31-42 (unused by DeepLabv3_plus).  Only `padding == dilation` (the 'same' case) is supported."""

    def __init__(self, inplanes, planes, kernel_size=3, stride=1, padding=0, dilation=1, bias=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, inplanes, kernel_size, stride, padding, dilation, groups=inplanes, bias=bias)
        self.pointwise = nn.Conv2d(inplanes, planes, 1, 1, 0, 1, 1, bias=bias)

    def _emit(self, eng, x, bn=None):
        if self.conv1.padding[0] != self.conv1.dilation[0]:
            raise NotImplementedError("deepcam_b200: SeparableConv2d needs padding == dilation")
        t = eng.dw(x, _dw_spec(self.conv1))
        return eng.conv(t, _conv_spec(self.pointwise), bn=bn)


def fixed_padding(inputs, kernel_size, rate):
    """DX:45-51.  Kept for API compatibility (plain tensor op); the engine folds this padding into the
    depthwise kernel's index arithmetic instead of materialising a padded copy."""
    k_eff = kernel_size + (kernel_size - 1) * (rate - 1)
    total = k_eff - 1
    beg = total // 2
    end = total - beg
    return torch.nn.functional.pad(inputs, (beg, end, beg, end))


class SeparableConv2d_same(_EngineModule):
    """DX:54-66: zero-pad by the dilation, depthwise 3x3 (stride, dilation), pointwise 1x1; no BN/ReLU between."""

    def __init__(self, inplanes, planes, kernel_size=3, stride=1, dilation=1, bias=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, inplanes, kernel_size, stride, 0, dilation, groups=inplanes, bias=bias)
        self.pointwise = nn.Conv2d(inplanes, planes, 1, 1, 0, 1, 1, bias=bias)

    def _emit(self, eng, x, bn=None, pre=None):
        """bn: BnSpec of the normalizer that follows (its batch sums then come out of the pointwise GEMM's epilogue).
        pre = (BnSpec, relu): x is still the PRE-BatchNorm tensor; the normalisation (+ReLU) in front of this unit is applied
        by the depthwise kernel while it loads its tile (Engine.bn_dw)."""
        if pre is not None:
            _, t = eng.bn_dw(x, pre[0], pre[1], _dw_spec(self.conv1))
        else:
            t = eng.dw(x, _dw_spec(self.conv1))
        return eng.conv(t, _conv_spec(self.pointwise), bn=bn)


# ---------------------------------------------------------------------------------------------------------
# DX:69-122
# ---------------------------------------------------------------------------------------------------------
class Block(_EngineModule):
    def __init__(self, inplanes, planes, reps, stride=1, dilation=1, start_with_relu=True, grow_first=True,
                 is_last=False, normalizer=nn.BatchNorm2d):
        super().__init__()
        if planes != inplanes or stride != 1:
            self.skip = nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False)
            self.skipbn = normalizer(planes)
        else:
            self.skip = None
        self.relu = nn.ReLU(inplace=True)
        self.start_with_relu = start_with_relu

        # (cin, cout) of the stride-1 separable units, in the reference's order (DX:82-97)
        units = []
        width = inplanes
        if grow_first:
            units.append((inplanes, planes))
            width = planes
        units += [(width, width)] * (reps - 1)
        if not grow_first:
            units.append((inplanes, planes))
        layers = []
        for cin, cout in units:
            layers += [self.relu, SeparableConv2d_same(cin, cout, 3, stride=1, dilation=dilation), normalizer(cout)]
        if not start_with_relu:
            layers = layers[1:]
        if stride != 1:
            layers.append(SeparableConv2d_same(planes, planes, 3, stride=2))
        if stride == 1 and is_last:
            layers.append(SeparableConv2d_same(planes, planes, 3, stride=1))
        self.rep = nn.Sequential(*layers)

    def _emit(self, eng, x, out_relu=False):
        """Returns the block output; `out_relu` asks for relu(output) to be stored (the consumer is a block whose
        leading in-place ReLU would clamp it anyway)."""
        inp = x
        if self.start_with_relu and not inp.is_relu:
            inp = eng.bn(inp, None, relu=True)               # explicit ReLU (stand-alone use on a raw tensor)
        skip_y = None
        if self.skip is not None:                             # emitted first so its strided dgrad accumulates last
            skip_y = eng.conv(inp, _conv_spec(self.skip), bn=_bn_spec(self.skipbn))
        cur, pending = inp, None
        mods = list(self.rep._modules.values())
        pre_relu = False
        for i, m in enumerate(mods):
            if isinstance(m, nn.ReLU):
                if pending is not None:
                    pre_relu = True                           # BatchNorm + this ReLU go into the next unit's depthwise load
                elif not cur.is_relu:
                    cur = eng.bn(cur, None, relu=True)
            elif isinstance(m, SeparableConv2d_same):
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                follows = nxt is not None and not isinstance(nxt, (nn.ReLU, SeparableConv2d_same))
                nbn = _bn_spec(nxt) if follows else None
                if pending is not None:
                    cur = m._emit(eng, pending[1], bn=nbn, pre=(pending[0], pre_relu))
                    pending, pre_relu = None, False
                else:
                    cur = m._emit(eng, cur, bn=nbn)
            else:
                pending = (_bn_spec(m), cur)
        if pending is not None and pre_relu:                  # (no layout of the reference ends in BatchNorm -> ReLU)
            cur, pending = eng.bn(pending[1], pending[0], relu=True), None
        if skip_y is not None:
            if pending is not None:
                cur = eng.bn(pending[1], pending[0], relu=False)
            return eng.bn(skip_y, _bn_spec(self.skipbn), relu=out_relu, residual=cur)
        if pending is not None:
            return eng.bn(pending[1], pending[0], relu=out_relu, residual=inp)
        return eng.bn(cur, None, relu=out_relu, residual=inp)


# ---------------------------------------------------------------------------------------------------------
# DX:125-280
# ---------------------------------------------------------------------------------------------------------
class Xception(_EngineModule):
    """Modified Aligned Xception (DX:125-242)."""

    def __init__(self, inplanes=3, os=16, pretrained=False, normalizer=nn.BatchNorm2d):
        super().__init__()
        if os == 16:
            entry_block3_stride, middle_rate, exit_rates = 2, 1, (1, 2)
        elif os == 8:
            entry_block3_stride, middle_rate, exit_rates = 1, 2, (2, 4)
        else:
            raise NotImplementedError

        self.conv1 = nn.Conv2d(inplanes, 32, 3, stride=2, padding=1, bias=False)
        self.bn1 = normalizer(32)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(32, 64, 3, stride=1, padding=1, bias=False)
        self.bn2 = normalizer(64)

        self.block1 = Block(64, 128, reps=2, stride=2, start_with_relu=False, normalizer=normalizer)
        self.block2 = Block(128, 256, reps=2, stride=2, start_with_relu=True, grow_first=True, normalizer=normalizer)
        self.block3 = Block(256, 728, reps=2, stride=entry_block3_stride, start_with_relu=True, grow_first=True,
                            is_last=True, normalizer=normalizer)
        for i in range(4, 20):                                # middle flow, DX:158-173
            setattr(self, "block%d" % i, Block(728, 728, reps=3, stride=1, dilation=middle_rate, start_with_relu=True,
                                               grow_first=True, normalizer=normalizer))
        self.block20 = Block(728, 1024, reps=2, stride=1, dilation=exit_rates[0], start_with_relu=True, grow_first=False,
                             is_last=True, normalizer=normalizer)
        self.conv3 = SeparableConv2d_same(1024, 1536, 3, stride=1, dilation=exit_rates[1])
        self.bn3 = normalizer(1536)
        self.conv4 = SeparableConv2d_same(1536, 1536, 3, stride=1, dilation=exit_rates[1])
        self.bn4 = normalizer(1536)
        self.conv5 = SeparableConv2d_same(1536, 2048, 3, stride=1, dilation=exit_rates[1])
        self.bn5 = normalizer(2048)

        self._init_weight()
        if pretrained:
            raise RuntimeError("deepcam_b200: pretrained ImageNet weights need network access (DX:254-280); "
                               "load a state_dict instead")

    def _init_weight(self):
        # DX:244-252: kaiming-normal on every Conv2d in module order, BatchNorm to (1, 0)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _emit(self, eng, x):
        h = eng.bn(eng.conv(x, _conv_spec(self.conv1), bn=_bn_spec(self.bn1)), _bn_spec(self.bn1), relu=True)
        h = eng.bn(eng.conv(h, _conv_spec(self.conv2), bn=_bn_spec(self.bn2)), _bn_spec(self.bn2), relu=True)
        h = self.block1._emit(eng, h, out_relu=True)
        low = h                                               # aliases relu(block1 output), DX:206 + SURVEY §0.3
        for i in range(2, 20):
            h = getattr(self, "block%d" % i)._emit(eng, h, out_relu=True)
        h = self.block20._emit(eng, h, out_relu=False)
        h = eng.bn(self.conv3._emit(eng, h, bn=_bn_spec(self.bn3)), _bn_spec(self.bn3), relu=True)
        h = eng.bn(self.conv4._emit(eng, h, bn=_bn_spec(self.bn4)), _bn_spec(self.bn4), relu=True)
        h = eng.bn(self.conv5._emit(eng, h, bn=_bn_spec(self.bn5)), _bn_spec(self.bn5), relu=True)
        return h, low

    def _emit_root(self, eng, x):
        h, low = self._emit(eng, x)
        return [(h, h.shape[3]), (low, low.shape[3])]

    def forward(self, x):
        h, low = _engine.run_module(self, [x])
        return h, low


# ---------------------------------------------------------------------------------------------------------
# DX:282-312
# ---------------------------------------------------------------------------------------------------------
class ASPP_module(_EngineModule):
    def __init__(self, inplanes, planes, rate, normalizer=nn.BatchNorm2d):
        super().__init__()
        kernel_size, padding = (1, 0) if rate == 1 else (3, rate)
        self.atrous_convolution = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, stride=1, padding=padding,
                                            dilation=rate, bias=False)
        self.bn = normalizer(planes)
        self.relu = nn.ReLU()
        torch.nn.init.kaiming_normal_(self.atrous_convolution.weight)      # DX:296, 304-312
        if isinstance(self.bn, nn.BatchNorm2d):
            self.bn.weight.data.fill_(1)
            self.bn.bias.data.zero_()

    def _emit(self, eng, x, out=None):
        y = eng.conv(x, _conv_spec(self.atrous_convolution), bn=_bn_spec(self.bn))
        return eng.bn(y, _bn_spec(self.bn), relu=True, out=out)


# ---------------------------------------------------------------------------------------------------------
# DX:315-395
# ---------------------------------------------------------------------------------------------------------
class InterpolationUpsampler(_EngineModule):
    """DX:315-344: the decoder variant the reference keeps beside DeconvUpsampler (DX:438 is commented out, so
    DeepLabv3_plus never builds it).  Same constructor, parameters and forward signature; the two
    F.interpolate(mode='bilinear', align_corners=True) calls run as dc_bilinear_fwd / dc_bilinear_bwd."""

    def __init__(self, n_output, normalizer=nn.BatchNorm2d):
        super().__init__()
        self.last_conv = nn.Sequential(nn.Conv2d(304, 256, kernel_size=3, stride=1, padding=1, bias=False),
                                       normalizer(256), nn.ReLU(),
                                       nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1, bias=False),
                                       normalizer(256), nn.ReLU(),
                                       nn.Conv2d(256, n_output, kernel_size=1, stride=1))
        self.n_output = n_output

    def _emit(self, eng, x, low):
        n = x.shape[0]
        hl, wl, cl = low.shape[1:]
        out_h, out_w = self._dc_plan_key
        if (hl, wl) != (int(math.ceil(out_h / 4)), int(math.ceil(out_w / 4))):
            # the reference fails in torch.cat (DX:329) for the same reason
            raise RuntimeError("InterpolationUpsampler: low_level_features are %dx%d, ceil(input_size / 4) is %dx%d"
                               % (hl, wl, int(math.ceil(out_h / 4)), int(math.ceil(out_w / 4))))
        c = x.shape[3]
        cat = eng.new_act(n, hl, wl, c + cl, x.t.dtype)
        eng.bn(low, None, relu=False, out=cat.slice(c, cl))                      # copy (torch.cat, DX:329)
        eng.bilinear(x, hl, wl, out=cat.slice(0, c))                             # DX:327-328
        lc = self.last_conv
        y = eng.bn(eng.conv(cat, _conv_spec(lc[0]), bn=_bn_spec(lc[1])), _bn_spec(lc[1]), relu=True)
        y = eng.bn(eng.conv(y, _conv_spec(lc[3]), bn=_bn_spec(lc[4])), _bn_spec(lc[4]), relu=True)
        pad = 8 if y.t.dtype == torch.bfloat16 else 4
        y = eng.conv(y, _conv_spec(lc[6]), out_c=(self.n_output + pad - 1) // pad * pad)
        return eng.bilinear(y, out_h, out_w, out_dtype=torch.float32)            # DX:331

    def _emit_root(self, eng, x, low):
        return [(self._emit(eng, x, low), self.n_output)]

    def forward(self, x, low_level_features, input_size):
        self._dc_plan_key = (int(input_size[-2]), int(input_size[-1]))
        return _engine.run_module(self, [x, low_level_features])[0]


class DeconvUpsampler(_EngineModule):
    def __init__(self, n_output, normalizer=nn.BatchNorm2d):
        super().__init__()

        def deconv(cout):
            return nn.ConvTranspose2d(256, cout, kernel_size=3, stride=2, padding=1, output_padding=(1, 1), bias=False)

        self.deconv1 = nn.Sequential(deconv(256), normalizer(256), nn.ReLU())
        self.deconv2 = nn.Sequential(deconv(256), normalizer(256), nn.ReLU())
        self.conv1 = nn.Sequential(nn.Conv2d(304, 256, kernel_size=3, stride=1, padding=1, bias=False),
                                   normalizer(256), nn.ReLU(),
                                   nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1, bias=False),
                                   normalizer(256), nn.ReLU(),
                                   nn.Conv2d(256, 256, kernel_size=1, stride=1))
        self.deconv3 = nn.Sequential(deconv(256), normalizer(256), nn.ReLU())
        self.last_deconv = nn.Sequential(deconv(n_output))
        self.n_output = n_output

    def _emit(self, eng, x, low=None, cat=None):
        """x: [N,h,w,256]; either `low` ([N,4h,4w,48], copied into the concat buffer) or `cat` (a [N,4h,4w,304]
        buffer whose channels 256..303 already hold the low-level features) must be given."""
        n, h, w, _ = x.shape
        y = eng.bn(eng.conv(x, _conv_spec(self.deconv1[0]), bn=_bn_spec(self.deconv1[1])), _bn_spec(self.deconv1[1]), relu=True)
        if cat is None:
            cat = eng.new_act(n, 4 * h, 4 * w, 256 + low.shape[3], x.t.dtype)
            eng.bn(low, None, relu=False, out=cat.slice(256, low.shape[3]))        # copy (torch.cat, DX:379)
        eng.bn(eng.conv(y, _conv_spec(self.deconv2[0]), bn=_bn_spec(self.deconv2[1])), _bn_spec(self.deconv2[1]), relu=True, out=cat.slice(0, 256))
        y = eng.bn(eng.conv(cat, _conv_spec(self.conv1[0]), bn=_bn_spec(self.conv1[1])), _bn_spec(self.conv1[1]), relu=True)
        y = eng.bn(eng.conv(y, _conv_spec(self.conv1[3]), bn=_bn_spec(self.conv1[4])), _bn_spec(self.conv1[4]), relu=True)
        y = eng.conv(y, _conv_spec(self.conv1[6]))
        y = eng.bn(eng.conv(y, _conv_spec(self.deconv3[0]), bn=_bn_spec(self.deconv3[1])), _bn_spec(self.deconv3[1]), relu=True)
        # logits stay fp32; channels padded for the vectorised kernels (8 in bf16 mode so that the logit gradient
        # is a legal TMA operand: 16-byte pixel pitch)
        co = self.n_output
        pad = 8 if y.t.dtype == torch.bfloat16 else 4
        return eng.conv(y, _conv_spec(self.last_deconv[0]), out_c=(co + pad - 1) // pad * pad, out_dtype=torch.float32)

    def _emit_root(self, eng, x, low):
        out = self._emit(eng, x, low=low)
        return [(out, self.n_output)]

    def forward(self, x, low_level_features, input_size=None):
        return _engine.run_module(self, [x, low_level_features])[0]


# ---------------------------------------------------------------------------------------------------------
# DX:398-480
# ---------------------------------------------------------------------------------------------------------
class DeepLabv3_plus(_EngineModule):
    def __init__(self, n_input=3, n_classes=21, os=16, pretrained=False, normalizer=nn.BatchNorm2d, _print=True, rank=0):
        if _print and (rank == 0):
            print("Constructing DeepLabv3+ model...")
            print("Number of output channels: {}".format(n_classes))
            print("Output stride: {}".format(os))
            print("Number of Input Channels: {}".format(n_input))
        super().__init__()
        self.xception_features = Xception(n_input, os, pretrained, normalizer)
        if os == 16:
            rates = [1, 6, 12, 18]
        elif os == 8:
            rates = [1, 12, 24, 36]
        else:
            raise NotImplementedError
        self.aspp1 = ASPP_module(2048, 256, rate=rates[0], normalizer=normalizer)
        self.aspp2 = ASPP_module(2048, 256, rate=rates[1], normalizer=normalizer)
        self.aspp3 = ASPP_module(2048, 256, rate=rates[2], normalizer=normalizer)
        self.aspp4 = ASPP_module(2048, 256, rate=rates[3], normalizer=normalizer)
        self.relu = nn.ReLU()
        self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)),
                                             nn.Conv2d(2048, 256, 1, stride=1, bias=False),
                                             normalizer(256), nn.ReLU())
        self.conv1 = nn.Conv2d(1280, 256, 1, bias=False)
        self.bn1 = normalizer(256)
        self.conv2 = nn.Conv2d(128, 48, 1, bias=False)       # [1x1, 48] channel reduction of the low-level features
        self.bn2 = normalizer(48)
        self.upsample = DeconvUpsampler(n_classes)
        self.n_classes = n_classes

    def _emit(self, eng, x):
        n, h, w, _ = x.shape
        if h % 16 or w % 16:
            raise RuntimeError("deepcam_b200: input height and width must be multiples of 16 (got %dx%d); the reference "
                               "fails with a concat size mismatch otherwise" % (h, w))
        feat, low = self.xception_features._emit(eng, x)
        fn, fh, fw, _ = feat.shape
        cat = eng.new_act(fn, fh, fw, 5 * 256, feat.t.dtype)                       # torch.cat target, DX:451
        # five independent branches over the same feature map, disjoint channel slices of the concat buffer (DX:443-451)
        for i, aspp in enumerate((self.aspp1, self.aspp2, self.aspp3, self.aspp4)):
            with eng.fork(i):
                aspp._emit(eng, feat, out=cat.slice(256 * i, 256))
        with eng.fork(4):
            g = eng.gap(feat)                                                      # fp32 [N,1,1,2048], DX:425
            g = eng.bn(eng.conv(g, _conv_spec(self.global_avg_pool[1]), bn=_bn_spec(self.global_avg_pool[2])), _bn_spec(self.global_avg_pool[2]), relu=True)
            eng.broadcast(g, cat.slice(1024, 256))                                 # DX:450
        eng.join()
        y = eng.bn(eng.conv(cat, _conv_spec(self.conv1), bn=_bn_spec(self.bn1)), _bn_spec(self.bn1), relu=True)
        ln, lh, lw, _ = low.shape
        dcat = eng.new_act(ln, lh, lw, 256 + 48, low.t.dtype)                      # torch.cat target, DX:379
        eng.bn(eng.conv(low, _conv_spec(self.conv2), bn=_bn_spec(self.bn2)), _bn_spec(self.bn2), relu=True, out=dcat.slice(256, 48))
        return self.upsample._emit(eng, y, cat=dcat)

    def _emit_root(self, eng, x):
        out = self._emit(eng, x)
        return [(out, self.n_classes)]

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()


def get_1x_lr_params(model):
    """DX:482-493."""
    for k in model.xception_features.parameters():
        if k.requires_grad:
            yield k


def get_10x_lr_params(model):
    """DX:496-505 (references `model.last_conv`, which the reference model does not have either)."""
    b = [model.aspp1, model.aspp2, model.aspp3, model.aspp4, model.conv1, model.conv2, model.last_conv]
    for j in range(len(b)):
        for k in b[j].parameters():
            if k.requires_grad:
                yield k
