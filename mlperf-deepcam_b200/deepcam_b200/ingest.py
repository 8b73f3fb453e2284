"""Input ingest on the device: the reference's CamDataset reads a sample as an [H, W, C] fp32 block,
transposes it to CHW on the host and normalises it with the per-channel statistics of stats.h5
(data/cam_hdf5_dataset.py:97-102, 122-129).  `normalize_hwc` does the normalisation on the GPU in the layout the file
already has and hands the network a channels-last tensor, so neither the host transpose nor the NCHW->NHWC transpose of
the model's first kernel is needed; the model (a drop-in torch.nn.Module) accepts the result like any [N, C, H, W] input.
"""
import ctypes

import torch

from . import _lib
from ._lib import DC_BF16, DC_F32, check


def stats_to_shift_scale(minval, maxval, device=None):
    """shift = minval, scale = 1 / (maxval - minval) as float32 (DS:99-102: computed in the statistics' dtype, then cast)."""
    minval = torch.as_tensor(minval)
    maxval = torch.as_tensor(maxval)
    shift = minval.to(torch.float32)
    scale = (1.0 / (maxval - minval)).to(torch.float32)
    if device is not None:
        shift, scale = shift.to(device), scale.to(device)
    return shift.contiguous(), scale.contiguous()


def normalize_hwc(raw, shift, scale, dtype=torch.bfloat16, out=None):
    """raw: CUDA float32 [N, H, W, C] (or [H, W, C]) in the file layout; shift/scale: CUDA float32 [C].
    Returns a tensor of logical shape [N, C, H, W] with channels-last strides holding (raw - shift) * scale in `dtype`
    (bfloat16: what the network computes in; float32: bit-identical to the reference's host expression)."""
    if not raw.is_cuda:
        raise RuntimeError("deepcam_b200.ingest: CUDA tensors required (there is no CPU fallback)")
    if raw.dim() == 3:
        raw = raw.unsqueeze(0)
    if raw.dim() != 4 or raw.dtype != torch.float32 or not raw.is_contiguous():
        raise ValueError("raw must be a contiguous float32 [N, H, W, C] tensor")
    n, h, w, c = raw.shape
    if c % 4:
        raise ValueError("channel count must be a multiple of 4")
    if dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("dtype must be float32 or bfloat16")
    for t in (shift, scale):
        if t.dtype != torch.float32 or t.numel() != c or not t.is_cuda or not t.is_contiguous():
            raise ValueError("shift/scale must be contiguous CUDA float32 vectors of %d elements" % c)
    if out is None:
        out = torch.empty((n, h, w, c), dtype=dtype, device=raw.device)
    elif out.shape != (n, h, w, c) or out.dtype != dtype or not out.is_contiguous():
        raise ValueError("out must be a contiguous [N, H, W, C] tensor of the requested dtype")
    stream = ctypes.c_void_p(torch.cuda.current_stream(raw.device).cuda_stream)
    rc = _lib.load().dc_ingest_hwc(ctypes.c_void_p(raw.data_ptr()), n * h * w, c, ctypes.c_void_p(shift.data_ptr()),
                                   ctypes.c_void_p(scale.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                   DC_BF16 if dtype == torch.bfloat16 else DC_F32, stream)
    check(rc, "dc_ingest_hwc")
    return out.permute(0, 3, 1, 2)
