"""stochy.jl_b200 -- B200-native particle inference for Stochy.jl's SMC / PMCMC / MH hot path.

The directory name carries a dot, so import it through the `stochy_jl_b200` shim at the repo root.
"""
from . import _native, dist
from ._native import build, lib
from .api import (BayesNet, Categorical, Discrete, DiscreteHMM, Engine, LinearGaussian, OpsProgram, StochyError, enum, hellingerdistance,
                  hmm2_example, hmm_example, kl, mh, pmcmc, score, smc, support)

__all__ = ["build", "lib", "Engine", "Discrete", "Categorical", "DiscreteHMM", "LinearGaussian", "OpsProgram", "BayesNet", "StochyError",
           "pmcmc", "smc", "mh", "enum", "hellingerdistance", "kl", "support", "score", "hmm2_example", "hmm_example"]
