"""Helpers that drive the CPU oracle (oracle/) from the tests.  Test infrastructure only."""
import ctypes as C

import numpy as np

from oracle import phys_ctypes as P


def dp(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """N oracle environments (oracle/cassie_env.c) stepped on the host."""

    def __init__(self, n, seed, dyn_rand, threads=8):
        self.L = P.lib()
        self.n, self.threads = n, threads
        self.buf = (C.c_char * (self.L.ce_sizeof_env() * n))()
        self.L.ce_batch_init(self.buf, n, C.c_uint(seed), int(dyn_rand), threads)
        self.obs = np.zeros((n, 50))
        self.rew = np.zeros(n)
        self.done = np.zeros(n, dtype=np.int32)
        self.term_obs = np.zeros((n, 50))

    def reset(self):
        self.L.ce_batch_reset(self.buf, self.n, dp(self.obs), self.threads)
        return self.obs

    def step(self, act, max_traj_len=400):
        act = np.ascontiguousarray(act, dtype=np.float64)
        self.L.ce_batch_step(self.buf, self.n, dp(act), dp(self.obs), dp(self.rew), dp(self.done), int(max_traj_len),
                             dp(self.term_obs), self.threads)
        return self.obs, self.rew, self.done
