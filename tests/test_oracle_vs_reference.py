"""Pins the oracle (oracle/deepcam_oracle.py) against the live reference classes.  Build container only."""
import os
import sys

import pytest
import torch

import refload

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import deepcam_oracle as O  # noqa: E402

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref_model():
    dx = refload.deeplab()
    torch.manual_seed(333)
    net = dx.DeepLabv3_plus(n_input=16, n_classes=3, os=16, pretrained=False, _print=False)
    return net


def test_init_state_dict_is_bit_identical(ref_model):
    sd = O.init_state_dict(16, 3, 16, seed=333)
    ref = ref_model.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert len(sd) == 532
    for k in ref:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
        assert torch.equal(sd[k], ref[k]), k


def test_forward_loss_backward_match_reference(ref_model):
    ls = refload.losses()
    sd = O.init_state_dict(16, 3, 16, seed=333)
    x, label = O.synthetic_batch(2, 64, 96, seed=7)
    w = O.class_weights()
    # reference
    ref_model.train()
    out_ref = ref_model(x.clone())
    loss_ref = ls.fp_loss(out_ref, label, weight=w, fpw_1=w[1], fpw_2=w[2])
    ref_model.zero_grad()
    loss_ref.backward()
    # oracle
    st = O.TrainState(sd)
    logits = O.forward(st.P, x.clone(), train=True)
    loss = O.fp_loss(logits, label, w, w[1], w[2])
    loss.backward()
    assert torch.allclose(logits, out_ref, rtol=0, atol=1e-5)
    assert abs(float(loss) - float(loss_ref)) < 1e-6
    ref_params = dict(ref_model.named_parameters())
    worst = 0.0
    for k in O.param_names(sd):
        g, gr = st.P[k].grad, ref_params[k].grad
        err = float((g - gr).norm() / (gr.norm() + 1e-20))
        worst = max(worst, err)
    assert worst < 1e-4, worst
    ref_sd = ref_model.state_dict()
    for k in ref_sd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert torch.allclose(st.P[k], ref_sd[k], rtol=1e-5, atol=1e-6), k
        if k.endswith("num_batches_tracked"):
            assert int(st.P[k]) == int(ref_sd[k]) == 1


def test_eval_forward_matches_reference(ref_model):
    sd = O.init_state_dict(16, 3, 16, seed=333)
    x, _ = O.synthetic_batch(1, 32, 48, seed=9)
    ref_model.load_state_dict(sd)
    ref_model.eval()
    with torch.no_grad():
        a = ref_model(x.clone())
        b = O.forward({k: v.clone() for k, v in sd.items()}, x.clone(), train=False)
    assert torch.allclose(a, b, rtol=0, atol=1e-5)
    ref_model.train()


def test_n1_training_raises_like_reference(ref_model):
    sd = O.init_state_dict(16, 3, 16, seed=333)
    x, _ = O.synthetic_batch(1, 32, 48, seed=9)
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        O.forward({k: v.clone() for k, v in sd.items()}, x, train=True)
    ref_model.train()
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        ref_model(x)


def test_fp_loss_and_score_known_answers():
    ls, ut = refload.losses(), refload.utils()
    torch.manual_seed(0)
    logit = torch.randn(2, 3, 8, 12)
    target = torch.randint(0, 3, (2, 8, 12))
    w = O.class_weights()
    a = ls.fp_loss(logit, target, weight=w, fpw_1=w[1], fpw_2=w[2])
    b = O.fp_loss(logit, target, w, w[1], w[2])
    assert abs(float(a) - float(b)) < 1e-6
    gt = torch.tensor([[0, 1, 1, 2], [1, 0, 0, 0]])
    pred = torch.tensor([[0, 1, 2, 2], [1, 1, 0, 0]])
    sa = ut.compute_score(pred, gt, num_classes=3, device_id=0)
    sb = O.compute_score(pred, gt, 3)
    assert float(sa) == float(sb)
    assert abs(float(sb) - 0.58333331) < 1e-6          # SURVEY §8c known answer
    z = torch.zeros(4, 4, dtype=torch.long)
    assert float(ut.compute_score(z, z, num_classes=3, device_id=0)) == float(O.compute_score(z, z, 3)) == 1.0
    for seed in range(3):
        g = torch.Generator().manual_seed(seed)
        p = torch.randint(0, 3, (2, 16, 24), generator=g)
        t = torch.randint(0, 3, (2, 16, 24), generator=g)
        assert float(ut.compute_score(p, t, num_classes=3, device_id=0)) == float(O.compute_score(p, t, 3))
