"""CPU-side checks: the C-ABI library builds/loads and exports every symbol declared in include/deepcam_b200.h
(no compute calls without a GPU), ctypes struct layouts match the header, and the tap tables of convdesc.py
express Conv2d / ConvTranspose2d fprop, dgrad and wgrad correctly (emulated gather-GEMM vs torch.nn.functional)."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn.functional as F

from deepcam_b200 import _lib, convdesc

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "deepcam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in _lib.SIGNATURES, "ctypes binding missing for " + n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.dc_abi_version() == _lib.DC_ABI_VERSION == 2
    assert isinstance(lib.dc_last_error_string(), bytes)


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.dc_view) == 8 + 4 * 4 + 4 * 8 + 8
    assert ctypes.sizeof(_lib.dc_conv_desc) == 4 * (1 + 3 * 9 + 4) + 4 + 4 + 8
    assert ctypes.sizeof(_lib.dc_bn_params) == 5 * 8 + 8 + 4 * 4
    assert _lib.dc_view.sn.offset == 24 and _lib.dc_view.dtype.offset == 56


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected on the host before any CUDA call; the error string names the problem."""
    lib = _lib.load()
    d = _lib.dc_conv_desc()
    d.ntaps = 0
    v = _lib.dc_view()
    rc = lib.dc_conv_gemm_simt(ctypes.byref(d), v, None, None, v, None)
    assert rc < 0
    assert b"ntaps" in lib.dc_last_error_string()
    d.ntaps, d.stride_h, d.stride_w, d.wtaps, d.out_csplit = 1, 1, 1, 1, 6      # two-segment outputs split at a multiple of 4
    rc = lib.dc_conv_gemm_simt(ctypes.byref(d), v, None, None, v, None)
    assert rc < 0 and b"out_csplit" in lib.dc_last_error_string()
    rc = lib.dc_bilinear_fwd(v, v, None)
    assert rc < 0 and b"dc_bilinear_fwd" in lib.dc_last_error_string()
    rc = lib.dc_pack_weight(ctypes.c_void_p(16), 8, 3, 9, 1, ctypes.c_void_p(16), _lib.DC_PACK_NTK_CONVT2, 64, 32, 1, None)
    assert rc < 0 and b"DC_PACK_NTK_CONVT2" in lib.dc_last_error_string()          # needs taps = 4
    rc = lib.dc_bn_stats(None, v, None)
    assert rc < 0 and b"dc_bn_stats" in lib.dc_last_error_string()
    assert lib.dc_bn_ws_bytes(728) == 32 * 728 + 64
    rc = lib.dc_iou_counts(None, None, 0, 3, None, None)
    assert rc < 0


# ---- gather-GEMM emulation ------------------------------------------------------------------------
def _gather(x, dh, dw, stride, ho, wo):
    """x: [N,H,W,C] -> [N,ho,wo,C] with x[n, y*s+dh, x*s+dw] or zero outside."""
    n, h, w, c = x.shape
    out = torch.zeros(n, ho, wo, c, dtype=x.dtype)
    for y in range(ho):
        iy = y * stride + dh
        if iy < 0 or iy >= h:
            continue
        for xx in range(wo):
            ix = xx * stride + dw
            if 0 <= ix < w:
                out[:, y, xx] = x[:, iy, ix]
    return out


def emul_gemm(taps, stride, x, wslices, ho, wo):
    """out[n,y,x,co] = sum_t gather_t(x)[..., ci] @ W[wt][ci][co]"""
    out = 0
    for dh, dw, wt in taps:
        out = out + _gather(x, dh, dw, stride, ho, wo) @ wslices[wt]
    return out


def emul_wgrad(taps, stride, x, dout, ntaps):
    """G[wt][co][ci] = sum_m gather_t(x)[m, ci] * dout[m, co]"""
    n, ho, wo, co = dout.shape
    G = torch.zeros(ntaps, co, x.shape[3], dtype=x.dtype)
    for dh, dw, wt in taps:
        g = _gather(x, dh, dw, stride, ho, wo)
        G[wt] += dout.reshape(-1, co).t() @ g.reshape(-1, x.shape[3])
    return G


@pytest.mark.parametrize("k,stride,pad,dil", [(1, 1, 0, 1), (3, 1, 1, 1), (3, 1, 6, 6), (3, 2, 1, 1), (1, 2, 0, 1)])
def test_conv2d_tap_tables(k, stride, pad, dil):
    torch.manual_seed(0)
    N, Ci, Co, H, W = 2, 5, 4, 9, 11
    x = torch.randn(N, Ci, H, W, dtype=torch.double, requires_grad=True)
    w = torch.randn(Co, Ci, k, k, dtype=torch.double, requires_grad=True)
    y = F.conv2d(x, w, None, stride, pad, dil)
    dy = torch.randn_like(y)
    y.backward(dy)
    ho, wo = y.shape[2:]
    assert ho == convdesc.conv_out_size(H, k, stride, pad, dil)
    xh, dyh = x.detach().permute(0, 2, 3, 1), dy.permute(0, 2, 3, 1)
    wf = [w.detach()[:, :, t // k, t % k].t() for t in range(k * k)]          # [ci][co] per tap
    out = emul_gemm(convdesc.conv_fprop_taps(k, pad, dil), stride, xh, wf, ho, wo)
    assert torch.allclose(out, y.detach().permute(0, 2, 3, 1), atol=1e-10)
    G = emul_wgrad(convdesc.conv_wgrad_taps(k, pad, dil), stride, xh, dyh, k * k)   # [tap][co][ci]
    assert torch.allclose(G.permute(1, 2, 0).reshape(Co, Ci, k, k), w.grad, atol=1e-9)
    if stride == 1:
        wd = [w.detach()[:, :, t // k, t % k] for t in range(k * k)]           # [co][ci] per tap (k = co)
        dx = emul_gemm(convdesc.conv_dgrad_taps(k, pad, dil), 1, dyh, wd, H, W)
        assert torch.allclose(dx, x.grad.permute(0, 2, 3, 1), atol=1e-10)


def test_conv_transpose_tap_tables():
    torch.manual_seed(1)
    N, Ci, Co, H, W = 2, 4, 3, 5, 6
    x = torch.randn(N, Ci, H, W, dtype=torch.double, requires_grad=True)
    w = torch.randn(Ci, Co, 3, 3, dtype=torch.double, requires_grad=True)
    y = F.conv_transpose2d(x, w, None, 2, 1, 1)
    assert y.shape[2:] == (2 * H, 2 * W)
    dy = torch.randn_like(y)
    y.backward(dy)
    xh, dyh = x.detach().permute(0, 2, 3, 1), dy.permute(0, 2, 3, 1)
    wf = [w.detach()[:, :, t // 3, t % 3] for t in range(9)]                  # [ci][co]
    out = torch.zeros(N, 2 * H, 2 * W, Co, dtype=torch.double)
    ntaps = 0
    for ph in range(2):
        for pw in range(2):
            taps = convdesc.convT_fprop_taps(3, 2, 1, ph, pw)
            ntaps += len(taps)
            out[:, ph::2, pw::2] = emul_gemm(taps, 1, xh, wf, H, W)
    assert ntaps == 9
    assert torch.allclose(out, y.detach().permute(0, 2, 3, 1), atol=1e-10)
    wd = [w.detach()[:, :, t // 3, t % 3].t() for t in range(9)]              # [co][ci] (k = co_out)
    dx = emul_gemm(convdesc.convT_dgrad_taps(3, 1), 2, dyh, wd, H, W)
    assert torch.allclose(dx, x.grad.permute(0, 2, 3, 1), atol=1e-10)
    # wgrad: gathered = dy (stride 2), enumerated = x  ->  G[tap][ci_in][co_out]
    G = emul_wgrad(convdesc.convT_wgrad_taps(3, 1), 2, dyh, xh, 9)
    assert torch.allclose(G.permute(1, 2, 0).reshape(Ci, Co, 3, 3), w.grad, atol=1e-9)


def test_conv_transpose_fused_tap_table_and_pack_definition():
    """One 2x2-tap contraction over all four output parities (convdesc.convT_fused_fprop_taps + the DC_PACK_NTK_CONVT2
    definition in include/deepcam_b200.h) + the two-segment output addressing == nn.ConvTranspose2d(k3, s2, p1, op1)."""
    torch.manual_seed(2)
    N, Ci, Co, H, W, G = 2, 4, 3, 5, 6, 4
    x = torch.randn(N, Ci, H, W, dtype=torch.double)
    w = torch.randn(Ci, Co, 3, 3, dtype=torch.double)
    y = F.conv_transpose2d(x, w, None, 2, 1, 1).permute(0, 2, 3, 1)
    wf = []
    for dh, dw, wt in convdesc.convT_fused_fprop_taps():
        assert wt == dh * 2 + dw
        m = torch.zeros(Ci, 4 * G, dtype=torch.double)                          # [ci][(a*2+b)*G + co]
        for a in range(2):
            for b in range(2):
                kh, kw = a + 1 - 2 * dh, b + 1 - 2 * dw
                if 0 <= kh <= 2 and 0 <= kw <= 2:
                    m[:, (a * 2 + b) * G:(a * 2 + b) * G + Co] = w[:, :, kh, kw]
        wf.append(m)
    fused = emul_gemm(convdesc.convT_fused_fprop_taps(), 1, x.permute(0, 2, 3, 1), wf, H, W)      # [N,H,W,4G]
    # two-segment store: channels < 2G at out[n, 2i, 2j] onwards (b, co contiguous), channels >= 2G one output row below
    out = torch.zeros(N, 2 * H, 2 * W, G, dtype=torch.double)
    flat = out.view(-1)
    sn, sh, sw, _ = out.stride()
    for n in range(N):
        for i in range(H):
            for j in range(W):
                base = n * sn + 2 * i * sh + 2 * j * sw
                for c in range(4 * G):
                    off = base + (sh + (c - 2 * G) if c >= 2 * G else c)
                    flat[off] = fused[n, i, j, c]
    assert torch.allclose(out[..., :Co], y, atol=1e-10)
    assert float(out[..., Co:].abs().max()) == 0.0


def _fake_view(n, h, w, c, dtype=_lib.DC_BF16, ptr=0x10000):
    """A dense NHWC view over a made-up, aligned address: the planning queries below never dereference it."""
    v = _lib.dc_view()
    v.ptr, v.n, v.h, v.w, v.c = ptr, n, h, w, c
    v.sc, v.sw, v.sh, v.sn = 1, c, w * c, h * w * c
    v.dtype = dtype
    return v


def _desc(taps, stride, wtaps):
    from deepcam_b200 import ops
    return ops.make_desc(taps, (stride, stride), False, wtaps)


def test_deterministic_workspace_planning_without_gpu(monkeypatch):
    """Host-side launch planning of the deterministic weight gradients (no kernel is launched, no device is touched): the workspace
    queries return splits x wtaps x Co x Ci for the pixel-split plan the launch would use - the per-tap kernel's one-wave split for
    the 728x728 layers, the halo-mode plan (all taps of a tap group per CTA) for conv1 / conv2 / last_deconv - and 0 when the
    launch does not split."""
    for k in ("DEEPCAM_B200_WGRAD_SPLIT_DIV", "DEEPCAM_B200_WGRAD_SPLIT_MUL", "DEEPCAM_B200_WGRAD_HALO", "DEEPCAM_B200_DWW_TILE",
              "DEEPCAM_B200_DW_S2_TILE", "DEEPCAM_B200_DW_TMA", "DEEPCAM_B200_DWW_BLOCKS_PER_SM", "DEEPCAM_B200_DWW_MIN_ROWS"):
        monkeypatch.delenv(k, raising=False)
    lib = _lib.load()
    q = lib.dc_conv_wgrad_tc_ws_elems
    # middle flow: 728 -> 728 pointwise at 2x48x72: 6 co tiles x 3 ci tiles of 256 -> 148 // 18 = 8 pixel splits
    d1 = _desc(convdesc.conv_wgrad_taps(1, 0, 1), 1, 1)
    assert q(ctypes.byref(d1), _fake_view(2, 48, 72, 728), _fake_view(2, 48, 72, 728)) == 8 * 728 * 728
    # ASPP 3x3 2048 -> 256 at 2x48x72: 2 x 8 x 9 = 144 tiles, no split, no workspace
    d9 = _desc(convdesc.conv_wgrad_taps(3, 6, 6), 1, 9)
    assert q(ctypes.byref(d9), _fake_view(2, 48, 72, 2048), _fake_view(2, 48, 72, 256)) == 0
    # conv1 16 -> 32 k3 s2 on 768x1152 (halo mode, one tap group): 3456 tiles of 8 x 16 pixels over 148 CTAs -> 24 per CTA -> 144 splits
    c1 = _desc(convdesc.conv_wgrad_taps(3, 1, 1), 2, 9)
    assert q(ctypes.byref(c1), _fake_view(2, 768, 1152, 16), _fake_view(2, 384, 576, 32)) == 144 * 9 * 32 * 16
    # conv2 32 -> 64 k3 s1 (halo mode, 9 taps x 32 channels = 288 > 192 -> two tap groups): 74 splits
    c2 = _desc(convdesc.conv_wgrad_taps(3, 1, 1), 1, 9)
    assert q(ctypes.byref(c2), _fake_view(2, 384, 576, 32), _fake_view(2, 384, 576, 64)) == 74 * 9 * 64 * 32
    # 64 gathered channels, 3 taps per group -> three groups; more than 64 channels -> per-tap kernel
    c3 = _desc(convdesc.conv_wgrad_taps(3, 1, 1), 1, 9)
    assert q(ctypes.byref(c3), _fake_view(2, 96, 144, 64), _fake_view(2, 96, 144, 128)) % (9 * 128 * 64) == 0
    bad = _fake_view(2, 48, 72, 730)                                    # C % 8 != 0: not a tcgen05 view
    assert q(ctypes.byref(d1), bad, _fake_view(2, 48, 72, 728)) == -1
    # depthwise: one slice of 9 x C per (x block, strip / image) - 9 x blocks x 10 (image, strip) pairs for the middle-flow tensors
    qd = lib.dc_dw_bwd_weight_ws_elems
    assert qd(_fake_view(2, 48, 72, 728), _fake_view(2, 48, 72, 728), 1, 1) == 9 * 10 * 9 * 728
    n_s2 = qd(_fake_view(2, 384, 576, 128), _fake_view(2, 192, 288, 128), 2, 1)
    assert n_s2 > 0 and n_s2 % (9 * 128) == 0
    assert qd(_fake_view(2, 48, 72, 728), _fake_view(2, 24, 36, 728), 1, 1) == -1          # output size does not match the stride
    # the switch itself is plain host state
    assert lib.dc_get_deterministic() == 0
    lib.dc_set_deterministic(1)
    assert lib.dc_get_deterministic() == 1
    lib.dc_set_deterministic(0)


def test_halo_mode_selection_without_gpu(monkeypatch):
    """dc_conv_gemm_tc_halo_ok is pure host planning: the halo mode of the persistent GEMM serves multi-tap gathers with <= 64
    gathered channels in dense NHWC rows (conv1, conv2, their gradients), not the 1x1 layers and not the wide 3x3 layers."""
    for k in ("DEEPCAM_B200_TC_HALO", "DEEPCAM_B200_TC_V1", "DEEPCAM_B200_TC_HALO_BUFS", "DEEPCAM_B200_TC_HALO_STAGES", "DEEPCAM_B200_TC_HALO_DBG"):
        monkeypatch.delenv(k, raising=False)
    lib = _lib.load()
    ok = lib.dc_conv_gemm_tc_halo_ok
    conv1 = _desc(convdesc.conv_fprop_taps(3, 1, 1), 2, 9)
    assert ok(ctypes.byref(conv1), _fake_view(2, 768, 1152, 16), _fake_view(2, 384, 576, 32)) == 1
    conv2 = _desc(convdesc.conv_fprop_taps(3, 1, 1), 1, 9)
    assert ok(ctypes.byref(conv2), _fake_view(2, 384, 576, 32), _fake_view(2, 384, 576, 64)) == 1
    assert ok(ctypes.byref(conv2), _fake_view(2, 192, 288, 256), _fake_view(2, 192, 288, 256)) == 0      # 256 gathered channels
    pw = _desc(convdesc.conv_fprop_taps(1, 0, 1), 1, 1)
    assert ok(ctypes.byref(pw), _fake_view(2, 48, 72, 64), _fake_view(2, 48, 72, 128)) == 0               # one tap: nothing to share
    strided = _fake_view(2, 384, 576, 32)
    strided.sw = 64                                                                                      # channel slice of a wider buffer
    assert ok(ctypes.byref(conv2), strided, _fake_view(2, 384, 576, 64)) == 0
