"""Device-side input ingest against the reference Dataset's host expression
(data/cam_hdf5_dataset.py:126-129: transpose HWC -> CHW, then data_scale * (data - data_shift))."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_getitem(raw_hwc, minval, maxval):
    """numpy restatement of CamDataset.__init__/__getitem__ normalisation (DS:97-102, 126-129) for one sample."""
    data_shift = minval
    data_scale = 1. / (maxval - data_shift)
    data_shift = np.reshape(data_shift, (data_shift.shape[0], 1, 1)).astype(np.float32)
    data_scale = np.reshape(data_scale, (data_scale.shape[0], 1, 1)).astype(np.float32)
    data = np.transpose(raw_hwc, (2, 0, 1))
    return data_scale * (data - data_shift)


@pytest.mark.parametrize("shape", [(2, 24, 40, 16), (1, 7, 9, 4), (3, 16, 16, 8)])
def test_normalize_hwc_matches_reference_dataset(shape):
    from deepcam_b200 import ingest
    rng = np.random.default_rng(4)
    n, h, w, c = shape
    raw = (rng.standard_normal(shape) * 50 + 200).astype(np.float32)
    minval = raw.reshape(-1, c).min(0).astype(np.float64) - 1.0      # stats.h5 holds float64 min / max
    maxval = raw.reshape(-1, c).max(0).astype(np.float64) + 1.0
    ref = np.stack([_reference_getitem(raw[i], minval, maxval) for i in range(n)])
    dev = torch.device("cuda:0")
    shift, scale = ingest.stats_to_shift_scale(minval, maxval, dev)
    x32 = ingest.normalize_hwc(torch.from_numpy(raw).to(dev), shift, scale, dtype=torch.float32)
    assert x32.shape == (n, c, h, w) and x32.stride(1) == 1              # logical NCHW, channels-last memory
    assert np.array_equal(x32.cpu().numpy(), ref)                        # bit-identical to the host expression
    x16 = ingest.normalize_hwc(torch.from_numpy(raw).to(dev), shift, scale, dtype=torch.bfloat16)
    assert torch.equal(x16.cpu(), torch.from_numpy(ref).to(torch.bfloat16))


def test_model_accepts_ingested_channels_last_input():
    """The drop-in module consumes the ingest output (bf16, channels-last) directly: same logits as the NCHW fp32 tensor
    holding the same (bf16-rounded) values."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import deepcam_oracle as O
    from architecture import deeplab_xception as dx
    from deepcam_b200 import ingest
    dev = torch.device("cuda:0")
    net = dx.DeepLabv3_plus(16, 3, 16, _print=False)
    net.load_state_dict(O.init_state_dict(16, 3, 16, seed=333))
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(8)
    raw = (torch.rand(2, 32, 48, 16, generator=g) * 300 + 100).to(dev)
    shift = torch.full((16,), 100.0, device=dev)
    scale = torch.full((16,), 1.0 / 300.0, device=dev)
    x_cl = ingest.normalize_hwc(raw, shift, scale, dtype=torch.bfloat16)
    x_nchw = x_cl.float().contiguous()
    with torch.no_grad():
        a = net(x_cl)
        b = net(x_nchw)
    assert torch.equal(a, b)
