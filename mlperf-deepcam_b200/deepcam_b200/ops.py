"""Tensor-level wrappers over the C ABI (include/deepcam_b200.h).

Activations are torch CUDA tensors in *logical* NHWC order, i.e. shape [N, H, W, C] with arbitrary strides
(a channel slice of a concat buffer, a parity sub-grid `t[:, ph::2, pw::2]`, or an NCHW tensor seen through
`.permute(0, 2, 3, 1)` are all valid).  torch is used for memory and streams only; every kernel is ours.
"""
import ctypes

import torch

from . import _lib
from ._lib import (DC_BF16, DC_BN_IDENTITY, DC_BN_RELU, DC_BN_RES_WRITE, DC_BN_TRAIN, DC_F32, DC_MAX_TAPS,
                   DC_PACK_NTK, DC_PACK_TKN, check, dc_bn_params, dc_conv_desc, dc_view)

_DT = {torch.float32: DC_F32, torch.bfloat16: DC_BF16}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("deepcam_b200 kernels need CUDA tensors (got %s); there is no CPU fallback" % t.device)


def view(t):
    """dc_view of a logical-NHWC tensor (None -> null view)."""
    if t is None:
        return dc_view()
    if t.dim() != 4:
        raise ValueError("expected a 4-D logical NHWC tensor, got shape %s" % (tuple(t.shape),))
    if t.dtype not in _DT:
        raise TypeError("unsupported dtype %s" % t.dtype)
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    # strides of size-1 dimensions are arbitrary in torch; normalise them so alignment checks are meaningful
    if c == 1:
        sc = 1
    if w == 1:
        sw = c * sc
    if h == 1:
        sh = w * sw
    if n == 1:
        sn = h * sh
    return dc_view(t.data_ptr(), n, h, w, c, sn, sh, sw, sc, _DT[t.dtype], 0)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


# ---- optional per-launch timing (bench.py roofline pass): CUDA events on the launching stream ------------
class Profiler:
    """Collects (kernel class, start event, end event, algorithmic FLOPs, algorithmic bytes) per launch."""

    def __init__(self, keep_launchers=False):
        self.items = []
        # one re-launchable closure per (kernel class, shape tag), with the tensors of its first occurrence kept alive:
        # graph_times() replays each of them back to back inside a CUDA graph, which is how the launch runs in the timed
        # region (the per-launch event pairs above add ~5-7 us of event/launch overhead to every short kernel)
        self.launchers = {} if keep_launchers else None

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, name, start, flops, nbytes, tag=""):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.items.append((name, start, e, flops, nbytes, tag))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, s, e, fl, nb, _tag in self.items:
            d = out.setdefault(name, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["flops"] += fl
            d["bytes"] += nb
        return out

    def graph_times(self, reps=10):
        """{class + ' ' + tag: microseconds per launch} with every recorded launch replayed `reps` times back to back inside
        one CUDA graph (warm-up replay first); kernels that cannot be captured are skipped."""
        out = {}
        if not self.launchers:
            return out
        torch.cuda.synchronize()
        for key, (fn, what) in self.launchers.items():
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(reps):
                        fn()
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                out[key[0] + " " + key[1]] = 1000.0 * e0.elapsed_time(e1) / reps
                del g
            except Exception:
                torch.cuda.synchronize()
        return out

    def by_tag(self, kinds):
        """Per-shape breakdown for the given kernel classes (call after summary(), which synchronises)."""
        out = {}
        for name, s, e, fl, nb, tag in self.items:
            if name not in kinds:
                continue
            d = out.setdefault(name + " " + tag, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["flops"] += fl
            d["bytes"] += nb
        return out


_prof = None


def set_profiler(p):
    global _prof
    _prof = p


def _nbytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def _shape_tag(t):
    return "%dx%dx%dx%d" % tuple(t.shape)


# Measurement aid (tools/ablate.py): kernel classes named in DEEPCAM_B200_ABLATE are NOT launched, so the step time
# difference is that class's share of the critical path.  Results are garbage while it is set; never set it otherwise.
import os as _os
_ABLATE = tuple(s for s in _os.environ.get("DEEPCAM_B200_ABLATE", "").split(",") if s)


def _timed(name, flops, nbytes, rc_fn, what, tag=""):
    if _ABLATE and any(name.startswith(a) for a in _ABLATE):
        return
    if _prof is None:
        check(rc_fn(), what)
        return
    st = _prof.begin()
    check(rc_fn(), what)
    _prof.end(name, st, flops, nbytes, tag)
    if _prof.launchers is not None and (name, tag) not in _prof.launchers:
        _prof.launchers[(name, tag)] = (rc_fn, what)


# ---- descriptors ---------------------------------------------------------------------------------
def make_desc(taps, stride=(1, 1), accumulate=False, wtaps=None, out_split=None, flags=0):
    """taps: sequence of (dh, dw, weight_slice).  out_split = (first channel of the second output segment, its element
    offset from the output view's origin), see dc_conv_desc.out_csplit."""
    if not 1 <= len(taps) <= DC_MAX_TAPS:
        raise ValueError("1..%d taps supported, got %d" % (DC_MAX_TAPS, len(taps)))
    d = dc_conv_desc()
    d.ntaps = len(taps)
    for i, (dh, dw, wt) in enumerate(taps):
        d.dh[i], d.dw[i], d.wt[i] = dh, dw, wt
    d.stride_h, d.stride_w = stride
    d.accumulate = 1 if accumulate else 0
    d.wtaps = wtaps if wtaps is not None else (max(t[2] for t in taps) + 1)
    if out_split is not None:
        d.out_csplit, d.out_split_off = int(out_split[0]), int(out_split[1])
    d.flags = int(flags)
    return d


# ---- layout / packing ------------------------------------------------------------------------------
def copy_view(src, dst):
    _require_cuda(src, dst)
    _timed("copy_view", 0.0, _nbytes(src, dst), lambda: _lib.load().dc_copy_view(view(src), view(dst), _stream()), "dc_copy_view")
    return dst


def fill_zero(t):
    _require_cuda(t)
    if not t.is_contiguous():
        raise ValueError("fill_zero needs a contiguous tensor")
    check(_lib.load().dc_fill_zero(_p(t), t.numel() * t.element_size(), _stream()), "dc_fill_zero")
    return t


def i64_increment_many(ptr_table, count):
    """ptr_table: int64 CUDA tensor holding `count` device addresses of int64 scalars."""
    check(_lib.load().dc_i64_increment_many(_p(ptr_table), count, _stream()), "dc_i64_increment_many")


def pack_weight(src, K, N, taps, src_k_first, layout, K_pad, N_pad, dtype, out=None):
    _require_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous()
    numel = taps * K_pad * N_pad
    if out is None:
        out = torch.empty(numel, dtype=dtype, device=src.device)
    assert out.numel() == numel and out.dtype == dtype
    check(_lib.load().dc_pack_weight(_p(src), K, N, taps, int(src_k_first), _p(out), layout, K_pad, N_pad, _DT[dtype], _stream()),
          "dc_pack_weight")
    return out


# blocks per pack job: every block first finds its job (a chain of dependent table reads), so a block must own enough work
# to amortise that; sweep knobs for tools/kbench.py --only pack
_PACK_MAX_BLOCKS = int(_os.environ.get("DEEPCAM_B200_PACK_MAX_BLOCKS", "128"))
_PACK_ELEMS_PER_BLOCK = int(_os.environ.get("DEEPCAM_B200_PACK_ELEMS_PER_BLOCK", "2048"))


def build_pack_table(jobs, device):
    """jobs: list of (src fp32 tensor, dst tensor, K, N, taps, src_k_first, layout, K_pad, N_pad).  Returns the device job
    table (uint8 tensor; keep it alive), the number of jobs and the grid size for pack_weights_multi."""
    arr = (_lib.dc_pack_job * len(jobs))()
    start = 0
    for i, (src, dst, K, N, taps, skf, layout, K_pad, N_pad) in enumerate(jobs):
        total = taps * K_pad * N_pad
        assert total < 2 ** 31 and src.dtype == torch.float32 and src.is_contiguous() and dst.numel() == total
        nb = max(1, min(_PACK_MAX_BLOCKS, (total + _PACK_ELEMS_PER_BLOCK - 1) // _PACK_ELEMS_PER_BLOCK))
        j = arr[i]
        j.src, j.dst = src.data_ptr(), dst.data_ptr()
        j.K, j.N, j.taps, j.src_k_first = K, N, taps, int(skf)
        j.layout, j.K_pad, j.N_pad, j.dst_dtype = layout, K_pad, N_pad, _DT[dst.dtype]
        j.block_start, j.n_blocks = start, nb
        start += nb
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device), len(jobs), start


def pack_weights_multi(table, njobs, total_blocks):
    _timed("pack_weights_multi", 0.0, 0.0,
           lambda: _lib.load().dc_pack_weights_multi(_p(table), njobs, total_blocks, _stream()), "dc_pack_weights_multi")


def unpack_wgrad(G, K, N, taps, dst_k_first, dst, k_stride=None):
    assert G.dtype == torch.float32 and dst.dtype == torch.float32 and dst.is_contiguous()
    assert dst.numel() == K * N * taps
    check(_lib.load().dc_unpack_wgrad(_p(G), K, N, taps, k_stride or K, int(dst_k_first), _p(dst), _stream()), "dc_unpack_wgrad")
    return dst


# ---- dense contractions -----------------------------------------------------------------------------
def conv_gemm(desc, x, w, bias, out, impl, bn_sums=None, flop_scale=1.0):
    """bn_sums (tcgen05 path only): zeroed BatchNorm workspace; the GEMM epilogue adds the batch sums of `out` to it.
    flop_scale: fraction of the launched MACs that are algorithmic (packs with structural zeros), for the profiler only."""
    _require_cuda(x, w, out)
    lib = _lib.load()
    m = out.shape[0] * out.shape[1] * out.shape[2]
    flops = 2.0 * m * out.shape[3] * x.shape[3] * desc.ntaps * flop_scale
    nbytes = _nbytes(x, out) + desc.ntaps * x.shape[3] * out.shape[3] * x.element_size() * flop_scale
    tag = "M%d Ci%d Co%d taps%d s%d" % (m, x.shape[3], out.shape[3], desc.ntaps, desc.stride_h)
    if desc.out_csplit:
        tag += " split%d" % desc.out_csplit
    if bn_sums is not None:
        assert impl == "tc" and bn_sums.dtype == torch.float64 and bn_sums.numel() >= bn_ws_elems(out.shape[3])
        _timed("conv_gemm_tc", flops, nbytes,
               lambda: lib.dc_conv_gemm_tc_bnstats(ctypes.byref(desc), view(x), _p(w), _p(bias), view(out), _p(bn_sums), _stream()),
               "dc_conv_gemm_tc_bnstats", tag=tag + " +bnstats")
        return out
    fn = lib.dc_conv_gemm_tc if impl == "tc" else lib.dc_conv_gemm_simt
    _timed("conv_gemm_" + impl, flops, nbytes,
           lambda: fn(ctypes.byref(desc), view(x), _p(w), _p(bias), view(out), _stream()), "dc_conv_gemm_" + impl, tag=tag)
    return out


def conv_halo_ok(desc, x, out):
    """True when the tcgen05 kernel would run this contraction in halo mode (dc_conv_gemm_tc_halo_ok): the weights must then
    be packed K-dense (K_pad = x.shape[3]) and the call flagged DC_CONV_HALO_PACK."""
    return _lib.load().dc_conv_gemm_tc_halo_ok(ctypes.byref(desc), view(x), view(out)) == 1


def conv_gemm_bn_eval(desc, x, w, bias, out, gamma, beta, running_mean, running_var, eps, relu):
    """out = [relu](bn_eval(conv(x) + bias)) in one tcgen05 launch (dc_conv_gemm_tc_bn_eval).  Returns False when the output
    layout needs the generic epilogue (nothing was launched)."""
    _require_cuda(x, w, out, gamma, beta, running_mean, running_var)
    lib = _lib.load()
    m = out.shape[0] * out.shape[1] * out.shape[2]
    flops = 2.0 * m * out.shape[3] * x.shape[3] * desc.ntaps
    nbytes = _nbytes(x, out) + desc.ntaps * x.shape[3] * out.shape[3] * x.element_size()
    state = {}

    def run():
        rc = lib.dc_conv_gemm_tc_bn_eval(ctypes.byref(desc), view(x), _p(w), _p(bias), view(out), _p(gamma), _p(beta),
                                         _p(running_mean), _p(running_var), float(eps), int(bool(relu)), _stream())
        state["rc"] = rc
        return 0 if rc == -2 else rc

    _timed("conv_gemm_tc", flops, nbytes, run, "dc_conv_gemm_tc_bn_eval",
           tag="M%d Ci%d Co%d taps%d s%d +bn_eval" % (m, x.shape[3], out.shape[3], desc.ntaps, desc.stride_h))
    return state["rc"] == 0


def set_deterministic(on):
    """Library switch for the kernels that need no workspace to be deterministic (the pooling reduction)."""
    lib = _lib.load()
    if bool(lib.dc_get_deterministic()) != bool(on):
        lib.dc_set_deterministic(int(bool(on)))


def conv_wgrad_ws_elems(desc, x, dout, impl):
    """fp32 elements of the workspace the deterministic two-stage weight gradient needs (0: the launch does not split)."""
    lib = _lib.load()
    fn = lib.dc_conv_wgrad_tc_ws_elems if impl == "tc" else lib.dc_conv_wgrad_simt_ws_elems
    n = fn(ctypes.byref(desc), view(x), view(dout))
    if n < 0:
        raise ValueError("dc_conv_wgrad_%s_ws_elems: bad arguments" % impl)
    return int(n)


def conv_wgrad(desc, x, dout, G, impl, ws=None):
    """ws (fp32, conv_wgrad_ws_elems() elements): deterministic two-stage form - per-split partial results in the workspace, added to
    G in split order by a second launch."""
    _require_cuda(x, dout, G)
    assert G.dtype == torch.float32
    lib = _lib.load()
    m = dout.shape[0] * dout.shape[1] * dout.shape[2]
    flops = 2.0 * m * dout.shape[3] * x.shape[3] * desc.ntaps
    nbytes = _nbytes(x, dout) + 4.0 * desc.ntaps * x.shape[3] * dout.shape[3]
    tag = "M%d Ci%d Co%d taps%d s%d" % (m, x.shape[3], dout.shape[3], desc.ntaps, desc.stride_h)
    if ws is not None:
        assert ws.dtype == torch.float32 and ws.is_contiguous()
        fn = lib.dc_conv_wgrad_tc_det if impl == "tc" else lib.dc_conv_wgrad_simt_det
        _timed("conv_wgrad_" + impl, flops, nbytes + 8.0 * ws.numel(),
               lambda: fn(ctypes.byref(desc), view(x), view(dout), _p(G), _p(ws), ws.numel(), _stream()), "dc_conv_wgrad_%s_det" % impl,
               tag=tag + " det")
        return G
    fn = lib.dc_conv_wgrad_tc if impl == "tc" else lib.dc_conv_wgrad_simt
    _timed("conv_wgrad_" + impl, flops, nbytes,
           lambda: fn(ctypes.byref(desc), view(x), view(dout), _p(G), _stream()), "dc_conv_wgrad_" + impl, tag=tag)
    return G


# ---- depthwise ----------------------------------------------------------------------------------------
def dw_fwd(x, w9c, stride, dil, out):
    _require_cuda(x, w9c, out)
    _timed("dw_fwd", 18.0 * out.numel(), _nbytes(x, out),
           lambda: _lib.load().dc_dw_fwd(view(x), _p(w9c), stride, dil, view(out), _stream()), "dc_dw_fwd",
           tag="%s s%d d%d" % (_shape_tag(x), stride, dil))
    return out


def dw_fwd_bn(params, y, w9c, act, out):
    """out = depthwise3x3(act), act = [relu](bn(y)) in one launch (dc_dw_fwd_bn); returns False when the tile does not fit
    (nothing was launched)."""
    _require_cuda(y, w9c, act, out)
    state = {}

    def run():
        rc = _lib.load().dc_dw_fwd_bn(ctypes.byref(params), view(y), _p(w9c), view(act), view(out), _stream())
        state["rc"] = rc
        return 0 if rc == -2 else rc

    _timed("dw_fwd_bn", 21.0 * out.numel(), _nbytes(y, act, out), run, "dc_dw_fwd_bn", tag="%s s1 d1" % _shape_tag(y))
    return state["rc"] == 0


def dw_bwd_data(dout, w9c, stride, dil, din, accumulate):
    _require_cuda(dout, w9c, din)
    _timed("dw_bwd_data", 18.0 * dout.numel(), _nbytes(dout, din) * (1.0 if not accumulate else 1.0) + (_nbytes(din) if accumulate else 0.0),
           lambda: _lib.load().dc_dw_bwd_data(view(dout), _p(w9c), stride, dil, view(din), int(accumulate), _stream()), "dc_dw_bwd_data",
           tag="%s s%d d%d acc%d" % (_shape_tag(din), stride, dil, int(accumulate)))
    return din


def dw_bwd_weight_ws_elems(x, dout, stride, dil):
    n = _lib.load().dc_dw_bwd_weight_ws_elems(view(x), view(dout), stride, dil)
    if n < 0:
        raise ValueError("dc_dw_bwd_weight_ws_elems: bad arguments")
    return int(n)


def dw_bwd_weight(x, dout, stride, dil, G9c, param_layout=False, ws=None):
    """ws (fp32, dw_bwd_weight_ws_elems() elements): deterministic two-stage form."""
    _require_cuda(x, dout, G9c)
    assert G9c.dtype == torch.float32
    if ws is not None:
        assert ws.dtype == torch.float32 and ws.is_contiguous()
        _timed("dw_bwd_weight", 18.0 * dout.numel(), _nbytes(x, dout),
               lambda: _lib.load().dc_dw_bwd_weight_det(view(x), view(dout), stride, dil, _p(G9c), int(param_layout), _p(ws), ws.numel(),
                                                        _stream()),
               "dc_dw_bwd_weight_det", tag="%s s%d d%d det" % (_shape_tag(x), stride, dil))
        return G9c
    _timed("dw_bwd_weight", 18.0 * dout.numel(), _nbytes(x, dout),
           lambda: _lib.load().dc_dw_bwd_weight(view(x), view(dout), stride, dil, _p(G9c), int(param_layout), _stream()),
           "dc_dw_bwd_weight",
           tag="%s s%d d%d" % (_shape_tag(x), stride, dil))
    return G9c


# ---- batch norm -----------------------------------------------------------------------------------------
def bn_ws_elems(c):
    """float64 elements of the per-layer BatchNorm workspace (dc_bn_ws_bytes)."""
    return (32 * c + 64) // 8


_onepass_cache = {}


def bn_onepass_ok(c, npix, dtype, backward):
    key = (c, npix, dtype, bool(backward))
    v = _onepass_cache.get(key)
    if v is None:
        v = _lib.load().dc_bn_onepass_ok(c, npix, _DT[dtype], int(bool(backward))) == 1
        _onepass_cache[key] = v
    return v


def bn_fwd_onepass(params, y, residual, out):
    _require_cuda(y, residual, out)
    _timed("bn_fwd_onepass", 6.0 * y.numel(), _nbytes(y, residual, out),
           lambda: _lib.load().dc_bn_fwd_onepass(ctypes.byref(params), view(y), view(residual), view(out), _stream()),
           "dc_bn_fwd_onepass", tag=_shape_tag(y) + (" res" if residual is not None else ""))
    return out


def bn_bwd_onepass(params, dout, out, y, rws, dy, dres, dgamma, dbeta):
    _require_cuda(dout, rws)
    _timed("bn_bwd_onepass", 12.0 * dout.numel(), _nbytes(dout, out, y, dy, dres),
           lambda: _lib.load().dc_bn_bwd_onepass(ctypes.byref(params), view(dout), view(out), view(y), _p(rws), view(dy),
                                                 view(dres), _p(dgamma), _p(dbeta), _stream()),
           "dc_bn_bwd_onepass", tag=_shape_tag(dout) + (" res" if dres is not None else ""))


def bn_stats(params, y):
    _require_cuda(y)
    _timed("bn_stats", 3.0 * y.numel(), _nbytes(y), lambda: _lib.load().dc_bn_stats(ctypes.byref(params), view(y), _stream()),
           "dc_bn_stats", tag=_shape_tag(y))


def bn_params(gamma, beta, running_mean, running_var, sums, count, momentum, eps, flags):
    p = dc_bn_params()
    p.gamma = gamma.data_ptr() if gamma is not None else None
    p.beta = beta.data_ptr() if beta is not None else None
    p.running_mean = running_mean.data_ptr() if running_mean is not None else None
    p.running_var = running_var.data_ptr() if running_var is not None else None
    p.sums = sums.data_ptr() if sums is not None else None
    p.count = float(count)
    p.momentum = float(momentum)
    p.eps = float(eps)
    p.flags = int(flags)
    return p


def bn_apply(params, y, residual, out):
    _require_cuda(y, residual, out)
    _timed("bn_apply", 3.0 * y.numel(), _nbytes(y, residual, out),
           lambda: _lib.load().dc_bn_apply(ctypes.byref(params), view(y), view(residual), view(out), _stream()), "dc_bn_apply",
           tag=_shape_tag(y) + (" res" if residual is not None else ""))
    return out


def bn_bwd_reduce(params, dout, out, y, rws, dgamma, dbeta):
    _require_cuda(dout, rws)
    assert rws.dtype == torch.float64 and rws.numel() >= bn_ws_elems(dout.shape[3])
    _timed("bn_bwd_reduce", 4.0 * dout.numel(), _nbytes(dout, out, y),
           lambda: _lib.load().dc_bn_bwd_reduce(ctypes.byref(params), view(dout), view(out), view(y), _p(rws), _p(dgamma),
                                                _p(dbeta), _stream()),
           "dc_bn_bwd_reduce", tag=_shape_tag(dout))
    return rws


def dw_bwd_data_bnred(dy, w9c, dx, accumulate, y, act, fwd_ws, bwd_ws, relu):
    """Depthwise backward-data fused with the ReLU mask and the BatchNorm backward reduction (dc_dw_bwd_data_bnred).
    Returns False when the kernel cannot take the shape (caller falls back to the separate kernels)."""
    _require_cuda(dy, w9c, dx, y)
    rc = [0]

    def run():
        rc[0] = _lib.load().dc_dw_bwd_data_bnred(view(dy), _p(w9c), view(dx), int(bool(accumulate)), view(y), view(act), _p(fwd_ws),
                                                 _p(bwd_ws), int(bool(relu)), _stream())
        return 0 if rc[0] == -2 else rc[0]
    _timed("dw_bwd_data_bnred", 22.0 * dx.numel(), _nbytes(dy, dx, y, act) + (dx.numel() * dx.element_size() if accumulate else 0), run,
           "dc_dw_bwd_data_bnred", tag=_shape_tag(dx) + (" res" if act is not None else ""))
    return rc[0] == 0


def bn_bwd_apply_finalize(params, dout, out, y, rws, dy, dres, dgamma, dbeta):
    _require_cuda(dout, y)
    _timed("bn_bwd_apply", 8.0 * dout.numel(), _nbytes(dout, out, y, dy, dres),
           lambda: _lib.load().dc_bn_bwd_apply_finalize(ctypes.byref(params), view(dout), view(out), view(y), _p(rws), view(dy),
                                                        view(dres), _p(dgamma), _p(dbeta), _stream()), "dc_bn_bwd_apply_finalize",
           tag=_shape_tag(dout) + " finalize" + (" res" if dres is not None else ""))


def bn_bwd_apply_reduced(params, g, y, rws, dy, dres, dgamma, dbeta):
    _require_cuda(g, y)
    _timed("bn_bwd_apply", 6.0 * g.numel(), _nbytes(g, y, dy, dres),
           lambda: _lib.load().dc_bn_bwd_apply_reduced(ctypes.byref(params), view(g), view(y), _p(rws), view(dy), view(dres),
                                                       _p(dgamma), _p(dbeta), _stream()), "dc_bn_bwd_apply_reduced",
           tag=_shape_tag(g) + " reduced" + (" res" if dres is not None else ""))


def bn_bwd_apply(params, dout, out, y, rws, dy, dres):
    _require_cuda(dout)
    _timed("bn_bwd_apply", 8.0 * dout.numel(), _nbytes(dout, out, y, dy, dres),
           lambda: _lib.load().dc_bn_bwd_apply(ctypes.byref(params), view(dout), view(out), view(y), _p(rws), view(dy),
                                               view(dres), _stream()), "dc_bn_bwd_apply",
           tag=_shape_tag(dout) + (" res" if dres is not None else "") + (" nody" if dy is None else ""))


def channel_sum(x, ws, out_c):
    _require_cuda(x, ws, out_c)
    assert ws.dtype == torch.float64 and ws.numel() >= x.shape[3] and out_c.dtype == torch.float32
    assert 1 <= out_c.numel() <= x.shape[3] and out_c.is_contiguous()
    check(_lib.load().dc_channel_sum(view(x), _p(ws), _p(out_c), out_c.numel(), _stream()), "dc_channel_sum")
    return out_c


# ---- pooling branch ---------------------------------------------------------------------------------------
def gap_fwd(x, mean_nc):
    _require_cuda(x, mean_nc)
    assert mean_nc.dtype == torch.float32 and mean_nc.is_contiguous()
    check(_lib.load().dc_gap_fwd(view(x), _p(mean_nc), _stream()), "dc_gap_fwd")
    return mean_nc


def reduce_hw(x, sum_nc):
    _require_cuda(x, sum_nc)
    assert sum_nc.dtype == torch.float32 and sum_nc.is_contiguous()
    check(_lib.load().dc_reduce_hw(view(x), _p(sum_nc), _stream()), "dc_reduce_hw")
    return sum_nc


def broadcast_hw(src_nc, dst):
    _require_cuda(src_nc, dst)
    assert src_nc.dtype == torch.float32 and src_nc.is_contiguous()
    check(_lib.load().dc_broadcast_hw(_p(src_nc), view(dst), _stream()), "dc_broadcast_hw")
    return dst


def gap_bwd(dmean_nc, dx, accumulate):
    _require_cuda(dmean_nc, dx)
    assert dmean_nc.dtype == torch.float32 and dmean_nc.is_contiguous()
    check(_lib.load().dc_gap_bwd(_p(dmean_nc), view(dx), int(accumulate), _stream()), "dc_gap_bwd")
    return dx


def bilinear_fwd(x, out):
    """out <- F.interpolate(x, size=out.shape[1:3], mode='bilinear', align_corners=True) on channels-last views (DX:327-331)."""
    _require_cuda(x, out)
    _timed("bilinear_fwd", 8.0 * out.numel(), _nbytes(x, out), lambda: _lib.load().dc_bilinear_fwd(view(x), view(out), _stream()),
           "dc_bilinear_fwd", tag=_shape_tag(out))
    return out


def bilinear_bwd(dout, din, accumulate):
    """din (+)= gradient of bilinear_fwd with respect to its input."""
    _require_cuda(dout, din)
    _timed("bilinear_bwd", 2.0 * 4 * dout.numel(), _nbytes(dout, din),
           lambda: _lib.load().dc_bilinear_bwd(view(dout), view(din), int(accumulate), _stream()), "dc_bilinear_bwd", tag=_shape_tag(dout))
    return din


# ---- loss / metric -------------------------------------------------------------------------------------------
def wce_fwd(logits_nhwc, target, class_w, acc, loss_out):
    _require_cuda(logits_nhwc, target, class_w)
    assert target.dtype == torch.int64 and target.is_contiguous()
    check(_lib.load().dc_wce_fwd(view(logits_nhwc), _p(target), _p(class_w), _p(acc), _p(loss_out), _stream()), "dc_wce_fwd")
    return loss_out


def wce_bwd(logits_nhwc, target, class_w, gscale, dlogits_nhwc):
    _require_cuda(logits_nhwc, target, class_w, dlogits_nhwc)
    assert target.dtype == torch.int64 and target.is_contiguous()
    check(_lib.load().dc_wce_bwd(view(logits_nhwc), _p(target), _p(class_w), _p(gscale), view(dlogits_nhwc), _stream()), "dc_wce_bwd")
    return dlogits_nhwc


def iou_counts(pred, gt, num_classes, counts):
    _require_cuda(pred, gt, counts)
    assert pred.dtype == torch.int64 and gt.dtype == torch.int64 and counts.dtype == torch.int64
    assert pred.is_contiguous() and gt.is_contiguous() and pred.numel() == gt.numel()
    check(_lib.load().dc_iou_counts(_p(pred), _p(gt), pred.numel(), num_classes, _p(counts), _stream()), "dc_iou_counts")
    return counts


def argmax_iou(logits_nhwc, gt, num_classes, pred_out, counts):
    _require_cuda(logits_nhwc)
    check(_lib.load().dc_argmax_iou(view(logits_nhwc), _p(gt), num_classes, _p(pred_out), _p(counts), _stream()), "dc_argmax_iou")


def iou_finalize(counts, num_classes, score_out):
    check(_lib.load().dc_iou_finalize(_p(counts), num_classes, _p(score_out), _stream()), "dc_iou_finalize")
    return score_out


def scale_f32(x, s):
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    check(_lib.load().dc_scale_f32(_p(x), x.numel(), float(s), _stream()), "dc_scale_f32")
    return x
