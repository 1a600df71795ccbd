"""apex_b200 — B200-native batched Cassie-v0 rollout + PPO update path (drop-in for osudrl/apex's
rl/algos/ppo.py sample()+update() over cassie/cassie.py), hand-written CUDA for sm_100a behind a C-ABI."""
from ._capi import lib, layout, ApexLibraryError  # noqa: F401

__all__ = ["lib", "layout", "ApexLibraryError"]
