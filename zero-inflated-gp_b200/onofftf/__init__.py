"""Host-side mirror of the reference's Kronecker (space (x) time) on/off model: the objective that scripts/onoff.py builds
from onofftf/main.py (``KernSE``, ``GaussKLkron``, ``DataSet``) -- same parameter names as that script -- evaluated by the
libzigp Kronecker context (csrc/kron.cuh, csrc/kron_host.cuh)."""
from .model import DataSet, OnOffKron
from .baselines import KronClassifier, KronSVGP
from . import utils

__all__ = ["OnOffKron", "DataSet", "KronSVGP", "KronClassifier", "utils"]
