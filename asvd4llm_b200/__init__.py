"""asvd4llm_b200 — B200-native (sm_100a) implementation of the ASVD hot path of hahnyuan/ASVD4LLM:
activation-scaled SVD factorisation of every nn.Linear and the SVDLinear low-rank forward, behind the
upstream Python operator surface.  See DESIGN.md."""
from . import _lib  # noqa: F401
from .modules.svd_linear import SVDLinear  # noqa: F401

__all__ = ["SVDLinear"]
