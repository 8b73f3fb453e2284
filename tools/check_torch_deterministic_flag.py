import os, sys, torch
sys.path.insert(0, "mlperf-deepcam_b200"); sys.path.insert(0, "oracle")
import deepcam_oracle as O
from architecture import deeplab_xception as dx
from utils import losses
from deepcam_b200 import engine
torch.use_deterministic_algorithms(True)
assert engine.deterministic()
sd = O.init_state_dict(16, 3, 16, seed=333); w = O.class_weights()
def run():
    net = dx.DeepLabv3_plus(16, 3, 16, _print=False); net.load_state_dict(sd); net.precision = "bf16"; net = net.to("cuda:0").train()
    res = []
    for i in range(3):
        x, label = O.synthetic_batch(2, 128, 192, seed=90 + i)
        net.zero_grad(); out = net(x.to("cuda:0")); loss = losses.fp_loss(out, label.to("cuda:0"), weight=w, fpw_1=w[1], fpw_2=w[2]); loss.backward()
        res.append([out.detach().clone(), loss.detach().clone()] + [p.grad.detach().clone() for p in net.parameters()])
    return res
a, b = run(), run()
bad = sum(1 for ra, rb in zip(a, b) for ta, tb in zip(ra, rb) if not torch.equal(ta, tb))
print("torch.use_deterministic_algorithms(True): mismatching tensors", bad, "of", sum(len(r) for r in a))
