"""Adapters that let tests drive the CUDA path (through the C ABI) like the oracle engines."""
import numpy as np
import torch

from ofighters_b200 import ArenaConfig, BatchedBattleground


class GpuEngine:
    def __init__(self, spawn, lcap=0, **kw):
        spawn = np.asarray(spawn)
        N, S, _ = spawn.shape
        self.bg = BatchedBattleground(N, ships={"external": S}, config=ArenaConfig(laser_cap=lcap),
                                      spawn_xy=torch.from_numpy(spawn.astype(np.int32)), **kw)

    def obs_vec(self):
        return self.bg.obs_vec.cpu().numpy()

    def step(self, actions):
        self.bg.generate_frame(torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int16)).to(self.bg.device))

    def reset(self, spawn):
        self.bg.restart(spawn_xy=torch.from_numpy(np.asarray(spawn, dtype=np.int32)).to(self.bg.device))

    def arrays(self):
        return {k: v.cpu().numpy() for k, v in self.bg.state().items()}

    def raster_bits(self):
        return self.bg.raster("bits").cpu().numpy().view(np.uint32)


class GpuEngineFused(GpuEngine):
    """Same protocol through the fused frame kernel (ofb_frame): every step also writes the observation maps."""

    def __init__(self, spawn, lcap=0, **kw):
        super().__init__(spawn, lcap=lcap, **kw)
        self.maps = self.bg.raster("bits")

    def step(self, actions):
        self.bg.generate_frame(torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int16)).to(self.bg.device),
                               maps=self.maps)

    def reset(self, spawn):
        super().reset(spawn)
        self.bg.raster("bits", out=self.maps)

    def raster_bits(self):
        return self.maps.cpu().numpy().view(np.uint32)
