"""TEST-ONLY interpreter of the engine's backend interface with plain torch ops on CPU.

Purpose: check the *graph logic* of deepcam_b200.engine and architecture/deeplab_xception.py (ReLU aliasing,
concat slices, gradient accumulation order, parameter-gradient routing, state_dict handling) against the
reference without a GPU.  It is never imported by the product package; the product backend is CUDA-only.
"""
import torch
import torch.nn.functional as F


def _nchw(t):
    return t.permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1)


class TorchBackend:
    name = "torch-test"

    def __init__(self, dtype=torch.float32, device=None):
        self.dtype = dtype
        self.device = torch.device("cpu")
        self.launches = 0
        self.log = []

    # ---- memory / layout ----
    def empty(self, n, h, w, c, dtype=None):
        return torch.full((n, h, w, c), float("nan"), dtype=dtype or self.dtype)

    def side_branch(self):
        import contextlib
        return contextlib.nullcontext()

    def fill_zero_flat(self, flat):
        flat.zero_()

    def increment_counters(self, counters):
        for c in counters:
            c += 1

    def from_nchw(self, x, c_pad=None):
        n, c, h, w = x.shape
        out = torch.zeros((n, h, w, c_pad or c), dtype=self.dtype)
        out[..., :c] = _nhwc(x).to(self.dtype)
        return out

    def to_nchw_f32(self, act, c):
        return _nchw(act[..., :c]).float().contiguous()

    # ---- dense conv ----
    def _f(self, t):
        return _nchw(t).double()

    def conv_fwd(self, x, spec, out, want_bn_sums=False):      # returns None: no fused statistics in the interpreter
        w = spec.weight.detach().double()
        b = spec.bias.detach().double() if spec.bias is not None else None
        if spec.transposed:
            y = F.conv_transpose2d(self._f(x), w, b, spec.stride, spec.pad, 1)
        else:
            y = F.conv2d(self._f(x), w, b, spec.stride, spec.pad, spec.dil)
        co = y.shape[1]
        out[..., :co] = _nhwc(y).to(out.dtype)
        if out.shape[3] > co:
            out[..., co:] = 0
        self.log.append(("conv_fwd", spec.name))
        return None if want_bn_sums else out

    def conv_bwd_data(self, dy, spec, dx, accumulate):
        w = spec.weight.detach().double()
        g = self._f(dy)[:, :spec.co]
        n, h, wd, c = dx.shape
        if spec.transposed:
            r = F.conv2d(g, w, None, spec.stride, spec.pad)
        else:
            r = torch.nn.grad.conv2d_input((n, c, h, wd), w, g, spec.stride, spec.pad, spec.dil)
        r = _nhwc(r)
        if accumulate:
            dx.copy_((dx.double() + r).to(dx.dtype))
        else:
            dx.copy_(r.to(dx.dtype))
        return dx

    def conv_bwd_weight(self, x, dy, spec, wgrad, bgrad=None):
        g = self._f(dy)[:, :spec.co]
        xx = self._f(x)
        with torch.enable_grad():
            w = spec.weight.detach().double().requires_grad_(True)
            if spec.transposed:
                y = F.conv_transpose2d(xx, w, None, spec.stride, spec.pad, 1)
            else:
                y = F.conv2d(xx, w, None, spec.stride, spec.pad, spec.dil)
            (gw,) = torch.autograd.grad(y, w, g)
        wgrad.copy_(gw.float())
        if bgrad is not None:
            bgrad.copy_(g.sum(dim=(0, 2, 3)).float())
        return wgrad

    # ---- depthwise ----
    def dw_fwd(self, x, spec, out):
        d = spec.dil
        y = F.conv2d(F.pad(self._f(x), (d, d, d, d)), spec.weight.detach().double(), None, spec.stride, 0, d, spec.c)
        out.copy_(_nhwc(y).to(out.dtype))
        return out

    def _dw_graph(self, x, spec):
        d = spec.dil
        with torch.enable_grad():
            xx = self._f(x).requires_grad_(True)
            w = spec.weight.detach().double().requires_grad_(True)
            y = F.conv2d(F.pad(xx, (d, d, d, d)), w, None, spec.stride, 0, d, spec.c)
        return xx, w, y

    def dw_bwd_data(self, dy, spec, dx, accumulate):
        xx, w, y = self._dw_graph(torch.zeros_like(dx), spec)
        (gx,) = torch.autograd.grad(y, xx, self._f(dy))
        r = _nhwc(gx)
        dx.copy_(((dx.double() + r) if accumulate else r).to(dx.dtype))
        return dx

    def dw_bwd_weight(self, x, dy, spec, wgrad):
        xx, w, y = self._dw_graph(x, spec)
        (gw,) = torch.autograd.grad(y, w, self._f(dy))
        wgrad.copy_(gw.float())
        return wgrad

    # ---- batch norm ----
    def bn_fwd(self, y, spec, relu, residual, out, training, ready_sums=None):
        v = y.double()
        saved = None
        if spec is not None:
            m = spec.module
            if training:
                cnt = v.numel() // v.shape[3]
                if cnt <= 1:
                    raise ValueError("Expected more than 1 value per channel when training, got input size %s"
                                     % (torch.Size((v.shape[0], v.shape[3], v.shape[1], v.shape[2])),))
                mean = v.mean(dim=(0, 1, 2))
                var = v.var(dim=(0, 1, 2), unbiased=False)
                if m.running_mean is not None:
                    mom = m.momentum if m.momentum is not None else 0.1
                    m.running_mean.mul_(1 - mom).add_(mom * mean.float())
                    m.running_var.mul_(1 - mom).add_(mom * (var * cnt / (cnt - 1)).float())
            else:
                mean, var = m.running_mean.double(), m.running_var.double()
            invstd = 1.0 / torch.sqrt(var + m.eps)
            saved = (mean, invstd)
            v = (v - mean) * invstd * m.weight.detach().double() + m.bias.detach().double()
        if residual is not None:
            v = v + residual.double()
        if relu:
            v = torch.relu(v)
        out.copy_(v.to(out.dtype))
        return saved

    def bn_bwd(self, dout, out, y, spec, sums, relu, dy, dres, res_accumulate, dgamma, dbeta, training=True):
        g = dout.double()
        if relu:
            g = g * (out.double() > 0)
        if dres is not None:
            dres.copy_(((dres.double() + g) if res_accumulate else g).to(dres.dtype))
        if spec is None:
            if dy is not None:
                dy.copy_(g.to(dy.dtype))
            return
        m = spec.module
        mean, invstd = sums
        xhat = (y.double() - mean) * invstd
        sg = g.sum(dim=(0, 1, 2))
        sgx = (g * xhat).sum(dim=(0, 1, 2))
        if dgamma is not None:
            dgamma.copy_(sgx.float())
        if dbeta is not None:
            dbeta.copy_(sg.float())
        if dy is not None:
            cnt = g.numel() // g.shape[3]
            scale = m.weight.detach().double() * invstd
            if training:
                r = scale * (g - sg / cnt - xhat * sgx / cnt)
            else:
                r = scale * g
            dy.copy_(r.to(dy.dtype))

    # ---- pooling branch ----
    def gap_fwd(self, x):
        return x.double().mean(dim=(1, 2)).float()

    def reduce_hw(self, x):
        return x.double().sum(dim=(1, 2)).float()

    def broadcast_hw(self, src_nc, out):
        out.copy_(src_nc[:, None, None, :].to(out.dtype).expand_as(out))
        return out

    def gap_bwd(self, dmean_nc, dx, accumulate):
        n, h, w, c = dx.shape
        r = (dmean_nc.double() / (h * w))[:, None, None, :].expand(n, h, w, c)
        dx.copy_(((dx.double() + r) if accumulate else r).to(dx.dtype))
        return dx


    def bilinear_fwd(self, x, out):
        y = F.interpolate(x.double().permute(0, 3, 1, 2), size=tuple(out.shape[1:3]), mode="bilinear", align_corners=True)
        out.copy_(y.permute(0, 2, 3, 1).to(out.dtype))
        return out

    def bilinear_bwd(self, dout, din, accumulate):
        n, h, w, c = din.shape
        with torch.enable_grad():
            x = torch.zeros(n, c, h, w, dtype=torch.double, requires_grad=True)
            y = F.interpolate(x, size=tuple(dout.shape[1:3]), mode="bilinear", align_corners=True)
            (g,) = torch.autograd.grad(y, x, dout.double().permute(0, 3, 1, 2))
        g = g.permute(0, 2, 3, 1)
        din.copy_(((din.double() + g) if accumulate else g).to(din.dtype))
        return din


def install():
    from deepcam_b200 import engine
    engine.set_backend_factory(lambda dtype, device: TorchBackend(dtype, device))


def uninstall():
    from deepcam_b200 import engine
    engine.set_backend_factory(None)
