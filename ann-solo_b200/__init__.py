"""ann-solo_b200 — B200-native hot path of ANN-SoLo's cascade open-modification search.

Import name: ``ann_solo_b200`` (the repository directory is ``ann-solo_b200/``; the importable
alias package at the repository root points here).
"""
from . import _lib
from ._lib import SoloError

__all__ = ["SoloError", "_lib"]
