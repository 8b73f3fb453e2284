"""Golden vectors generated from the LIVE reference (tests/golden/make_golden.py) — they travel to the GPU box.
CPU: the oracle reproduces them.  GPU: the product path (fp32 mode) reproduces them through the C-ABI kernels."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import deepcam_oracle as O  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "deepcam_ref_small.npz"))


def _rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double().reshape(-1)
    b = torch.as_tensor(np.asarray(b)).double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _kept():
    return [k[len("grad::"):] for k in G.files if k.startswith("grad::")]


def test_oracle_reproduces_reference_golden_vectors():
    sd = O.init_state_dict(16, 3, 16, seed=333)
    names = O.param_names(sd)
    assert names == list(G["param_names"])
    assert np.allclose([float(sd[k].double().sum()) for k in names], G["param_checksums"], rtol=0, atol=1e-9)
    x, label = O.synthetic_batch(2, 32, 48, seed=2024)
    st = O.TrainState(sd)
    logits = O.forward(st.P, x, train=True)
    w = O.class_weights()
    loss = O.fp_loss(logits, label, w, w[1], w[2])
    loss.backward()
    assert _rel(logits.detach(), G["logits"]) < 1e-5
    assert abs(float(loss) - float(G["loss"])) < 1e-6
    pred = torch.max(logits, 1)[1]
    assert np.array_equal(pred.numpy().astype(np.int8), G["pred"])
    assert sum(O.confusion_counts(pred, label, 3), []) == G["confusion"].tolist()
    assert float(O.compute_score(pred, label, 3)) == float(G["score"])
    norms = np.array([float(st.P[k].grad.double().norm()) for k in names])
    assert np.allclose(norms, G["grad_norms"], rtol=2e-3)
    for k in _kept():
        assert _rel(st.P[k].grad.reshape(-1)[:4096], G["grad::" + k]) < 1e-3, k
    rm = [float(v.double().sum()) for k, v in st.P.items() if k.endswith("running_mean")]
    assert np.allclose(rm, G["running_mean_sums"], rtol=1e-4, atol=1e-5)
    ev = O.forward({k: v.clone() for k, v in sd.items()}, x[:1], train=False)
    assert _rel(ev, G["logits_eval"]) < 1e-5
    torch.manual_seed(0)
    lg, tg = torch.randn(2, 3, 8, 12), torch.randint(0, 3, (2, 8, 12))
    assert abs(float(O.fp_loss(lg, tg, w)) - float(G["fp_loss_seed0"])) < 1e-6
    assert abs(float(G["fp_loss_seed0"]) - 2.3611667) < 1e-5


@pytest.mark.gpu
def test_product_path_reproduces_reference_golden_vectors():
    from architecture import deeplab_xception as dx
    from utils import losses, utils as dcutils
    dev = "cuda:0"
    torch.manual_seed(333)
    net = dx.DeepLabv3_plus(16, 3, 16, _print=False)            # same seed, same constructor order as the reference
    names = [k for k, _ in net.named_parameters()]
    assert names == list(G["param_names"])
    assert np.allclose([float(p.double().sum()) for p in net.parameters()], G["param_checksums"], rtol=0, atol=1e-9)
    net.precision = "fp32"
    net = net.to(dev).train()
    x, label = O.synthetic_batch(2, 32, 48, seed=2024)
    w = O.class_weights()
    out = net(x.to(dev))
    loss = losses.fp_loss(out, label.to(dev), weight=w, fpw_1=w[1], fpw_2=w[2])
    loss.backward()
    assert _rel(out.detach().cpu(), G["logits"]) < 2e-4
    assert abs(float(loss) - float(G["loss"])) < 1e-5
    pred = torch.max(out, 1)[1]
    agree = float((pred.cpu().numpy().astype(np.int8) == G["pred"]).mean())
    assert agree > 0.999
    if agree == 1.0:
        assert dcutils.iou_counts(pred, label.to(dev), 3).cpu().tolist() == G["confusion"].tolist()
        assert float(dcutils.compute_score(pred, label.to(dev), num_classes=3, device_id=0)) == float(G["score"])
    # identical predictions -> identical integer counters, independent of the forward pass
    gp = torch.from_numpy(G["pred"].astype(np.int64)).to(dev)
    assert dcutils.iou_counts(gp, label.to(dev), 3).cpu().tolist() == G["confusion"].tolist()
    assert float(dcutils.compute_score(gp, label.to(dev), num_classes=3, device_id=0)) == float(G["score"])
    norms = np.array([float(p.grad.double().norm()) for p in net.parameters()])
    assert np.allclose(norms, G["grad_norms"], rtol=5e-2)
    net.eval()
    net.load_state_dict(O.init_state_dict(16, 3, 16, seed=333))
    with torch.no_grad():
        ev = net(x[:1].to(dev))
    assert _rel(ev.cpu(), G["logits_eval"]) < 1e-4
