"""OnOffLikelihood with the reference's surface (onoffgpf/OnOffLikelihood.py:24-32): a noise variance
``Param(0.01, positive)`` and ``variational_expectations(Fmu, Fvar, Fmuvar, Y)``.  Inside the model the same expression
is evaluated per point by csrc/nstage.cuh:point_eval; this method is the stand-alone API mirror."""
from __future__ import annotations

import math

import torch

from .param import DTYPE, Param, Parameterized, positive


class OnOffLikelihood(Parameterized):
    def __init__(self):
        self.variance = Param(0.01, positive)

    def variational_expectations(self, Fmu, Fvar, Fmuvar, Y):
        v = self.variance.constrained()
        Fmu, Fvar, Fmuvar, Y = (torch.as_tensor(t, dtype=DTYPE, device=v.device) for t in (Fmu, Fvar, Fmuvar, Y))
        return -0.5 * math.log(2 * math.pi) - 0.5 * torch.log(v) - 0.5 * (torch.square(Y - Fmu) + Fvar + Fmuvar) / v
