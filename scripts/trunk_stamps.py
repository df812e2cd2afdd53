"""Phase timeline of CTA 0 of k_st_trunk12 (clock64): python scripts/trunk_stamps.py [frames] [n_arenas]"""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from ofighters_b200 import BatchedBattleground, _lib
from ofighters_b200.policy import PolicyB200

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = int(sys.argv[2]) if len(sys.argv) > 2 else 296 * 8
bg = BatchedBattleground(n, ships={"random": 7}, seed=5)
for _ in range(frames):
    bg.frame()
maps = bg.raster("bits")
vec = bg.obs_vec[:, 0, :].contiguous()
pol = PolicyB200.random_init(device=bg.device, seed=0, max_ships=n)
lib = _lib.load()
lib.ofb_policy_st_stamps.argtypes = [ctypes.c_void_p]
pol.forward(maps, vec, want_act=False, want_argmax=False)
torch.cuda.synchronize()
st = torch.zeros(128 + 8 * 16, dtype=torch.int64, device=bg.device)
lib.ofb_policy_st_stamps(ctypes.c_void_p(st.data_ptr()))
pol.forward(maps, vec, want_act=False, want_argmax=False)
torch.cuda.synchronize()
lib.ofb_policy_st_stamps(None)
s = st.cpu()[:128].reshape(8, 16)
t = st.cpu()[128:].reshape(8, 16)
print("arena | fill   mapwait  Q      D/prefix lists  conv1  level2 level3 level4 end   | total   n1    n2    n3    n4")
for i in range(1, 7):
    r = [int(x) for x in s[i]]
    print("%5d | %6d %6d %6d %6d %6d %6d %6d %6d %6d %6d | %6d %5d %5d %5d %5d" % (
        i, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[7] - r[6], r[12] - r[7], r[13] - r[12],
        r[14] - r[13], r[14] - r[0], r[8], r[9], r[10], r[11]))
print("first tile of a level, seen by thread 0 (a draining thread): level start -> its MMAs done / drain (cycles)")
for i in range(1, 7):
    q = [int(x) for x in t[i]]
    r = [int(x) for x in s[i]]
    print("%5d | L2 %5d %4d | L3 %5d %4d | L4 %5d %4d" % (i, q[2] - r[6], q[3] - q[2], q[6] - r[7], q[7] - q[6], q[10] - r[12], q[11] - q[10]))
