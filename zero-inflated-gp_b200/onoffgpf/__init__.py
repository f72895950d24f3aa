"""Host-side mirror of the reference package ``onoffgpf`` (same import name, same public names:
``from onoffgpf import OnOffSVGP, OnOffLikelihood``, reference onoffgpf/__init__.py:1-4), plus the minimal ``RBF`` kernel
and ``Param`` objects that the reference takes from GPflow 0.4."""
from . import kernels
from .kernels import RBF
from .likelihood import OnOffLikelihood
from .model import OnOffSVGP
from .param import Param, positive
from ._cabi import NotPositiveDefiniteError, ZigpError

__all__ = ["OnOffSVGP", "OnOffLikelihood", "RBF", "kernels", "Param", "positive", "ZigpError", "NotPositiveDefiniteError"]
