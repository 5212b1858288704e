"""Dual-averaging step-size tuner for HMC (Hoffman & Gelman 2014, algorithms 4-5); mirror of
eeyore/tuners/hmcda_tuner.py:8-59.

With the native HMC the recurrence runs on the device, one independent tuner per chain, inside the fused sampler kernel
(eeyore_b200/csrc/samplers.cuh:da_tune); this object carries the hyper-parameters and, for a single chain, mirrors the
state after a run.  `tune` keeps the reference's host implementation for API compatibility.
"""
import numpy as np

from .tuner import Tuner


class HMCDATuner(Tuner):
    def __init__(self, l, e0=None, d=0.65, eub=None):
        self.l, self.e0, self.d, self.eub = l, e0, d, eub
        self.m = None if e0 is None else np.log(10 * e0)
        self.logeub = None if eub is None else np.log(eub)
        self.logbare, self.barh = 0.0, 0.0
        self.g, self.t0, self.k = 0.05, 10, 0.75

    def set_m(self, e0):
        self.m = np.log(10 * e0)

    def num_steps(self, e):
        return max(1, round(self.l / e))

    def tune(self, rate, idx, return_e=True):
        it = idx + 1
        d_w, e_w = 1 / (it + self.t0), 1 / (it ** self.k)
        self.barh = (1 - d_w) * self.barh + d_w * (self.d - rate)
        loge = self.m - np.sqrt(it) * self.barh / self.g
        if self.logeub is not None:
            loge = min(loge, self.logeub)
        self.logbare = e_w * loge + (1 - e_w) * self.logbare
        e = np.exp(loge) if return_e else np.exp(self.logbare)
        return e, self.num_steps(e)
