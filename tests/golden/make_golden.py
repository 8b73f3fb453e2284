"""Generates tests/golden/deepcam_ref_small.npz from the LIVE reference (/root/reference, build container only).

The reference ships no golden vectors, so these are outputs of the reference classes themselves:
  model  = DeepLabv3_plus(n_input=16, n_classes=3, os=16) constructed under torch.manual_seed(333)   (DX:398-439)
  input  = oracle.synthetic_batch(2, 32, 48, seed=2024) (uniform [0,1) tiles, class-frequency labels)
  loss   = fp_loss(logits, label, class weights TR:204-209)                                           (LS:28-52)
  score  = compute_score(argmax, label, 3)                                                            (UT:32-60)
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle")]
import refload  # noqa: E402
import deepcam_oracle as O  # noqa: E402


def main():
    dx, ls, ut = refload.deeplab(), refload.losses(), refload.utils()
    torch.manual_seed(333)
    net = dx.DeepLabv3_plus(n_input=16, n_classes=3, os=16, pretrained=False, _print=False)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    x, label = O.synthetic_batch(2, 32, 48, seed=2024)
    w = O.class_weights()
    net.train()
    out = net(x.clone())
    loss = ls.fp_loss(out, label, weight=w, fpw_1=w[1], fpw_2=w[2])
    net.zero_grad()
    loss.backward()
    names = [k for k, _ in net.named_parameters()]
    grads = dict(net.named_parameters())
    pred = torch.max(out, 1)[1]
    score = ut.compute_score(pred, label, num_classes=3, device_id=0)
    sd1 = net.state_dict()
    # eval-mode forward with the *initial* buffers
    net2 = dx.DeepLabv3_plus(n_input=16, n_classes=3, os=16, pretrained=False, _print=False)
    net2.load_state_dict(sd0)
    net2.eval()
    with torch.no_grad():
        out_eval = net2(x[:1].clone())
    keep = ["xception_features.conv1.weight", "xception_features.block1.skip.weight", "xception_features.block4.rep.1.conv1.weight",
            "xception_features.block10.rep.4.pointwise.weight", "aspp3.bn.weight", "global_avg_pool.1.weight", "conv2.weight",
            "upsample.conv1.6.bias", "upsample.last_deconv.0.weight"]
    data = dict(
        logits=out.detach().numpy(), loss=np.float32(loss.item()), score=np.float32(score.item()),
        pred=pred.numpy().astype(np.int8), logits_eval=out_eval.numpy(),
        param_names=np.array(names), grad_norms=np.array([float(grads[k].grad.double().norm()) for k in names]),
        param_checksums=np.array([float(sd0[k].double().sum()) for k in names]),
        running_mean_sums=np.array([float(v.double().sum()) for k, v in sd1.items() if k.endswith("running_mean")]),
        running_var_sums=np.array([float(v.double().sum()) for k, v in sd1.items() if k.endswith("running_var")]),
        confusion=np.array(sum(O.confusion_counts(pred, label, 3), []), dtype=np.int64),
        fp_loss_seed0=np.float32(ls.fp_loss(torch.manual_seed(0) and torch.randn(2, 3, 8, 12), torch.randint(0, 3, (2, 8, 12)),
                                            weight=w, fpw_1=w[1], fpw_2=w[2]).item()),
    )
    for k in keep:
        data["grad::" + k] = grads[k].grad.reshape(-1)[:4096].numpy()          # leading 4096 elements keep the file small
    path = os.path.join(HERE, "deepcam_ref_small.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(loss), "score", float(score))


if __name__ == "__main__":
    main()
