import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_train import _arena_batch, _dense_image
from oracle import policy_torch as po
from oracle import policy_train_torch as pt
from ofighters_b200.trainer import TrainerB200
B = 8
w = po.init_weights(4, randomize_bn=True)
_, maps, vec = _arena_batch(B)
g = torch.Generator().manual_seed(B)
ta = torch.randn((B, 2), generator=g) * 3
tp = torch.randn((B, 400, 400), generator=g) * 0.5
tr = TrainerB200(weights=w, learning_rate=1e-4, batch_size=8)
img, vec_h = _dense_image(maps), vec.cpu()
(olosses, ograds, _) = pt.loss_and_grads(w, img, vec_h, ta, tp)
wd = {k: v.double() for k, v in w.items()}
(_, dgrads, _) = pt.loss_and_grads(wd, img.double(), vec_h.double(), ta.double(), tp.double())
loss = tr.fit(maps, vec, ta.cuda(), tp.cuda()).cpu()
print("loss", loss.tolist(), olosses)
grads = tr.get_grads()
for k, og in ograds.items():
    dg = dgrads[k].float()
    print("%-18s max|g64| %.3e  gpu-vs-64 %.3e  torch32-vs-64 %.3e" % (k, float(dg.abs().max()), float((grads[k] - dg).abs().max()), float((og - dg).abs().max())))
