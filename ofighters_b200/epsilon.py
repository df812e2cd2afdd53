"""Exploration schedules of the shared trainer -- the B200 side of ``Trainer.epsilon`` (agents/qlearnIA_V2.py:53,195-201).

The reference keeps epsilon as mutable host state advanced by ``next()`` (lib/epsilon.py:36-86).  Here a schedule is a
closed-form function eps(t) of the number t of ``decay_epsilon()`` calls, because the batched eps-greedy branch runs inside
a kernel (``ofb_policy_play_actions`` evaluates the same formula on the device from the ``ofb_eps_schedule`` struct and t).
``get() / next() / set()`` keep the reference's calling convention so a schedule object drops into ``Trainer(epsilon=...)``:

    reference                      here
    -----------------------------  ----------------------------------------------
    Epsilon_cos(period)            EpsilonSchedule.cosine(period)     (alias Epsilon_cos)
    Epsilon_decay()                EpsilonSchedule.decaying()         (alias Epsilon_decay)
    eps.get() / .next() / .set(v)  same names; .t is the schedule step, .value_at(t) the closed form
"""
import math

from . import _lib


class EpsilonSchedule:
    def __init__(self, kind, start=1.0, period=0.0, decay=1.0, floor=0.0):
        if kind not in (_lib.EPS_CONST, _lib.EPS_COSINE, _lib.EPS_DECAY):
            raise Exception("unknown epsilon schedule kind " + str(kind))
        if kind == _lib.EPS_COSINE and not period > 0:
            raise Exception("a cosine schedule needs a positive period")
        self.kind, self.start, self.period, self.decay, self.floor = kind, float(start), float(period), float(decay), float(floor)
        self.t = 0.0

    # ---- constructors
    @classmethod
    def constant(cls, value):
        return cls(_lib.EPS_CONST, start=value)

    @classmethod
    def cosine(cls, period, amplitude=1.0):
        """eps(t) = amplitude * (cos(2 pi t / period) + 1) / 2: starts at `amplitude`, 0 at half a period, periodic."""
        return cls(_lib.EPS_COSINE, start=amplitude, period=period)

    @classmethod
    def decaying(cls, start=1.0, decay=0.99990, floor=0.01):
        """eps(t) = start * decay^t until it has reached `floor`, constant from then on (Epsilon_decay's constants by default)."""
        return cls(_lib.EPS_DECAY, start=start, decay=decay, floor=floor)

    # ---- closed form (the device evaluates the same expression, ofb_policy.cu: eps_at)
    def _stop(self):
        if self.start > self.floor and 0.0 < self.decay < 1.0:
            return math.ceil(math.log(self.floor / self.start) / math.log(self.decay))
        return 0.0

    def value_at(self, t):
        if self.kind == _lib.EPS_COSINE:
            return self.start * (math.cos(2.0 * math.pi * math.fmod(t, self.period) / self.period) + 1.0) * 0.5
        if self.kind == _lib.EPS_DECAY:
            return self.start * self.decay ** min(float(t), self._stop())
        return self.start

    # ---- reference calling convention
    def get(self):
        return self.value_at(self.t)

    def next(self):
        self.t += 1.0
        if self.kind == _lib.EPS_COSINE and self.t >= self.period:
            self.t -= self.period
        return self.get()

    def set(self, value):
        """Jump to the point of the schedule whose epsilon is `value` (first half-period of a cosine; a decaying or constant
        schedule restarts from `value`)."""
        if not 0.0 <= value <= 1.0:
            raise Exception("epsilon must lie in [0, 1], got {}".format(value))
        if self.kind == _lib.EPS_COSINE:
            if value > self.start:
                raise Exception("epsilon {} is above the schedule's amplitude {}".format(value, self.start))
            self.t = self.period * math.acos(2.0 * value / self.start - 1.0) / (2.0 * math.pi)
        else:
            self.start, self.t = float(value), 0.0

    def as_struct(self):
        s = _lib.OfbEpsSchedule()
        s.kind, s.start, s.period, s.decay, s.floor = self.kind, self.start, self.period, self.decay, self.floor
        return s


def Epsilon_cos(period):
    """Drop-in name of lib/epsilon.py:36."""
    return EpsilonSchedule.cosine(period)


def Epsilon_decay():
    """Drop-in name of lib/epsilon.py:62."""
    return EpsilonSchedule.decaying()
