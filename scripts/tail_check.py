"""Bring-up check of the fused tail kernel (k_tz_tail): against the two-kernel tail (same operands' arithmetic) and the
fp32 oracle, with the error split by image region so that a wrong border variant shows up where it lives."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import policy_torch as po
from ofighters_b200 import BatchedBattleground
from ofighters_b200.policy import PolicyB200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
bil = sys.argv[2] if len(sys.argv) > 2 else "tf2"
w = po.init_weights(5, randomize_bn=True)
bg = BatchedBattleground(n, ships={"random": 7}, seed=11)
for _ in range(30):
    bg.frame()
maps = bg.raster("bits")
vec = bg.obs_vec[:, 0, :].contiguous()
fused = PolicyB200(w, max_ships=max(16, n), bilinear=bil)
fused.set_taps(True)
unf = PolicyB200(w, max_ships=max(16, n), bilinear=bil, fused_tail=False)
rf = fused.forward(maps, vec, 1, want_ptr=True)
ru = unf.forward(maps, vec, 1, want_ptr=True)
torch.cuda.synchronize()
u3f = fused.debug_tap(6, n, (200, 200, 8)).float().cpu()
u3u = unf.debug_tap(6, n, (200, 200, 8)).float().cpu()
u2f = fused.debug_tap(5, n, (100, 100, 8)).float().cpu()
u2u = unf.debug_tap(5, n, (100, 100, 8)).float().cpu()
print("up2 pairs layout == plane layout:", bool(torch.equal(u2f, u2u)))


def regions(d, name):
    d = d.abs()
    H = d.shape[1]
    inner = d[:, 1:H - 1, 1:H - 1]
    print("%-10s max %.3e | interior %.3e top %.3e bottom %.3e left %.3e right %.3e | corners %s" % (
        name, float(d.max()), float(inner.max()), float(d[:, 0].max()), float(d[:, H - 1].max()), float(d[:, :, 0].max()),
        float(d[:, :, H - 1].max()), ["%.2e" % float(d[:, y, x].max()) for y in (0, H - 1) for x in (0, H - 1)]))
    if float(d.max()) > 0:
        k = int(torch.argmax(d.reshape(-1)))
        print("           worst at", np.unravel_index(k, d.shape))


print("scale up3 %.3f ptr %.3f" % (float(u3u.abs().max()), float(ru["ptr"].abs().max())))
regions(u3f - u3u, "up3 f-u")
regions((rf["ptr"] - ru["ptr"]).cpu(), "ptr f-u")
print("xy fused == unfused:", bool(torch.equal(rf["xy"], ru["xy"])), rf["xy"][:4].cpu().tolist())
k = torch.argmax(rf["ptr"].reshape(n, -1), dim=1).cpu()
print("xy == argmax(own dense map):", bool(torch.equal(rf["xy"].cpu().long(), torch.stack([k % 400, k // 400], dim=1))))
b = maps.cpu().numpy().view(np.uint32)
img = np.unpackbits(b.view(np.uint8).reshape(n, 2, -1), axis=2, bitorder="little").reshape(n, 2, 400, 400)
act, ptr, inter = po.forward(w, torch.from_numpy(img.transpose(0, 2, 3, 1).astype(np.float32)), vec.cpu(), return_intermediates=True,
                             **({"bilinear": bil} if bil != "tf2" else {}))
regions(u3f - inter["up3"].permute(0, 2, 3, 1), "up3 f-orc")
regions(u3u - inter["up3"].permute(0, 2, 3, 1), "up3 u-orc")
regions(rf["ptr"].cpu() - ptr, "ptr f-orc")
regions(ru["ptr"].cpu() - ptr, "ptr u-orc")
# forward_argmax without the dense map (the product path) gives the same xy
i2, xy2 = fused.forward_argmax(maps, vec, 1)
print("argmax-only path == with dense map:", bool(torch.equal(xy2, rf["xy"])))
# ---- where do fused and two-kernel maps differ?
d = (rf["ptr"] - ru["ptr"]).abs().cpu()
tol = 0.05
rows = (d.amax(dim=(0, 2)) > tol).nonzero().flatten().tolist()
cols = (d.amax(dim=(0, 1)) > tol).nonzero().flatten().tolist()
print("rows V with |diff| > %.2f: %s" % (tol, rows[:60]), len(rows))
print("cols Z with |diff| > %.2f: %s" % (tol, cols[:60]), len(cols))
bad = (d > tol)
print("bad pixels per ship (first 16):", bad.sum(dim=(1, 2)).tolist()[:16])
ys, xs = bad[0].nonzero(as_tuple=True)
print("ship 0 bad (V, Z):", list(zip(ys.tolist(), xs.tolist()))[:80])
for (V, Z) in list(zip(ys.tolist(), xs.tolist()))[:6]:
    print("  (%d,%d): fused %.4f unfused %.4f oracle %.4f" % (V, Z, float(rf["ptr"][0, V, Z]), float(ru["ptr"][0, V, Z]), float(ptr[0, V, Z])))
