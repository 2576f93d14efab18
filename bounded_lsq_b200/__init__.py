"""B200-native trust-region inner loop of bounded nonlinear least squares.

Public surface mirrors ``bounded_lsq/__init__.py:3-13`` for the trust-region
path: ``least_squares``, ``trf``, ``dogbox`` and the bound helpers, plus the
batched entry point ``least_squares_batched``.
"""
from .least_squares import (least_squares, least_squares_batched, trf, dogbox,
                            stage_host_inputs, OptimizeResult, TERMINATION_MESSAGES)
from .batched import PerProblem
from ._lib import BlsqError, get_lib

__all__ = ["least_squares", "least_squares_batched", "trf", "dogbox", "stage_host_inputs",
           "PerProblem", "OptimizeResult", "TERMINATION_MESSAGES",
           "BlsqError", "get_lib"]
