"""adcraft_b200: B200-native batched AdCraft simulator (the BiddingSimulation.step hot path).

Only the hot path lives here: CUDA kernels + C ABI (csrc/, include/adcraft_b200.h) and the
host-side mirror of the reference's env interface.  Importing the package does not need a GPU;
constructing an env does (there is no CPU fallback).
"""
from .keywords import (EXPLICIT, IMPLICIT, KeywordTable, sample_implicit_keywords_from_quantiles,
                       sample_random_keywords)

__all__ = [
    "EXPLICIT", "IMPLICIT", "KeywordTable", "sample_implicit_keywords_from_quantiles",
    "sample_random_keywords", "VectorBiddingSimulation",
]
__version__ = "0.1.0"


def __getattr__(name):  # torch is imported lazily, with the env classes
    if name == "VectorBiddingSimulation":
        from .vector_env import VectorBiddingSimulation
        return VectorBiddingSimulation
    raise AttributeError(name)
