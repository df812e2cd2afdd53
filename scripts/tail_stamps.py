"""Timeline of CTA 0 of the fused tail kernel (clock64 stamps per tile): python scripts/tail_stamps.py [n_ships]
columns per tile g: drain3 {acc ready, packed, slot free, stored}, mma4 {ring full, tmem free}, drain4 {acc ready, done}"""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from ofighters_b200 import BatchedBattleground, _lib
from ofighters_b200.policy import PolicyB200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 4
bg = BatchedBattleground(n, ships={"random": 7}, seed=5)
for _ in range(30):
    bg.frame()
maps = bg.raster("bits")
vec = bg.obs_vec[:, 0, :].contiguous()
pol = PolicyB200.random_init(device=bg.device, seed=0, max_ships=n)
lib = _lib.load()
lib.ofb_policy_tail_stamps.argtypes = [ctypes.c_void_p]
pol.forward_argmax(maps, vec)
torch.cuda.synchronize()
st = torch.zeros(64 * 8, dtype=torch.int64, device=bg.device)
lib.ofb_policy_tail_stamps(ctypes.c_void_p(st.data_ptr()))
pol.forward_argmax(maps, vec)
torch.cuda.synchronize()
lib.ofb_policy_tail_stamps(None)
s = st.cpu().reshape(64, 8)
t0 = int(s[0, 0])
print("tile | d3: acc   packed  free    stored | m4: ringfull tmemfree | d4: acc   done   (cycles since tile 0's accumulator)")
for g in range(48):
    print("%4d | %7d %7d %7d %7d | %7d %7d | %7d %7d" % ((g,) + tuple(int(x) - t0 for x in s[g])))
d = (s[21:42, 7] - s[20:41, 7]).float()
print("steady state: %.0f cycles per tile (drain4 done to done), min %.0f max %.0f" % (float(d.mean()), float(d.min()), float(d.max())))
