"""LR / weight-decay schedules written into ``optimizer.param_groups`` once per step (host).

Drop-in for the reference's ``src/utils/schedulers.py`` (``WarmupCosineSchedule :11-45``,
``CosineWDSchedule :48-76``): same constructor arguments, ``step()`` returns the new value,
groups flagged ``WD_exclude`` keep their weight decay.
"""
import math


class WarmupCosineSchedule(object):
    """Linear warm-up start_lr -> ref_lr, then half-cosine ref_lr -> final_lr."""

    def __init__(self, optimizer, warmup_steps, start_lr, ref_lr, T_max, last_epoch=-1, final_lr=0.):
        self.optimizer = optimizer
        self.start_lr = start_lr
        self.ref_lr = ref_lr
        self.final_lr = final_lr
        self.warmup_steps = warmup_steps
        self.T_max = T_max - warmup_steps     # length of the cosine leg
        self._step = 0.

    def value(self, step):
        if step < self.warmup_steps:
            frac = float(step) / float(max(1, self.warmup_steps))
            return self.start_lr + frac * (self.ref_lr - self.start_lr)
        frac = float(step - self.warmup_steps) / float(max(1, self.T_max))
        cos_lr = self.final_lr + (self.ref_lr - self.final_lr) * 0.5 * (1. + math.cos(math.pi * frac))
        return max(self.final_lr, cos_lr)

    def step(self):
        self._step += 1
        new_lr = self.value(self._step)
        for group in self.optimizer.param_groups:
            group['lr'] = new_lr
        return new_lr


class CosineWDSchedule(object):
    """Half-cosine ref_wd -> final_wd over T_max steps, clamped on the final_wd side."""

    def __init__(self, optimizer, ref_wd, T_max, final_wd=0.):
        self.optimizer = optimizer
        self.ref_wd = ref_wd
        self.final_wd = final_wd
        self.T_max = T_max
        self._step = 0.

    def value(self, step):
        wd = self.final_wd + (self.ref_wd - self.final_wd) * 0.5 * (1. + math.cos(math.pi * step / self.T_max))
        return max(self.final_wd, wd) if self.final_wd <= self.ref_wd else min(self.final_wd, wd)

    def step(self):
        self._step += 1
        new_wd = self.value(self._step)
        for group in self.optimizer.param_groups:
            if not group.get('WD_exclude', False):
                group['weight_decay'] = new_wd
        return new_wd
