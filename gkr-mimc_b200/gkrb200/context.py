"""Device context: arena + stream + (optional) multi-GPU communicator.  Replaces poly/pool.go and
sumcheck/worker.go of the reference."""
import ctypes

import numpy as np

from ._lib import Stats, check, lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def fr_array(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] != 4:
        raise ValueError("field elements must have a trailing dimension of 4 uint64 limbs")
    return a


def fr_empty(*shape):
    return np.zeros(tuple(shape) + (4,), dtype=np.uint64)


class Context:
    def __init__(self, device=0, max_bn=16, stream=None, world=1):
        """world > 1: the context will join a communicator of that many ranks; its arena is sized for the 1/world shard"""
        self._h = ctypes.c_void_p()
        check(lib().gkrb200_init_shard(ctypes.byref(self._h), device, max_bn, ctypes.c_void_p(stream) if stream else None, world))
        self.device, self.max_bn = device, max_bn
        self.rank, self.world = 0, 1

    # -- multi-GPU ---------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (ctypes.c_uint8 * 128)()
        check(lib().gkrb200_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, world, unique_id, leader=0):
        """leader: the rank that runs the transcript of this communicator's proofs (same value on every rank)"""
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        check(lib().gkrb200_comm_init(self._h, rank, world, buf))
        self.rank, self.world = rank, world
        if world > 1:
            check(lib().gkrb200_comm_set_leader(self._h, leader))

    @property
    def exchange_mode(self):
        """'window' (round kernels publish into host-shared mapped memory), 'nccl' (all-gather) or None (single GPU)"""
        return {0: "window", 1: "nccl"}.get(lib().gkrb200_comm_exchange_mode(self._h))

    # -- lifetime ----------------------------------------------------------------------------
    def close(self):
        if self._h:
            lib().gkrb200_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # -- instrumentation ---------------------------------------------------------------------
    def stats_reset(self):
        check(lib().gkrb200_stats_reset(self._h))

    def stats(self):
        s = Stats()
        check(lib().gkrb200_stats_get(self._h, ctypes.byref(s)))
        return s

    def set_profiling(self, on):
        check(lib().gkrb200_set_profiling(self._h, 1 if on else 0))

    OPT_GENERIC_CIPHER, OPT_PAR8_MAX_PAIRS, OPT_HOST_TAIL_LEN, OPT_CF_BLOCKS_PER_SM, OPT_EXCHANGE, OPT_INLINE_MIN_PAIRS, OPT_TRANSCRIPT, OPT_CONST_FOLD = 1, 2, 3, 4, 5, 6, 7, 8

    def set_option(self, option, value):
        check(lib().gkrb200_set_option(self._h, option, int(value)))

    def microbench(self, kind, iters):
        rate, ms = ctypes.c_double(), ctypes.c_double()
        check(lib().gkrb200_microbench(self._h, kind, iters, ctypes.byref(rate), ctypes.byref(ms)))
        return rate.value, ms.value

    # -- device field ops (parity tests) -----------------------------------------------------
    def fr_batch(self, op, a, b=None):
        a = fr_array(a).reshape(-1, 4)
        b = fr_array(b).reshape(-1, 4) if b is not None else None
        out = fr_empty(a.shape[0])
        check(lib().gkrb200_fr_batch(self._h, op, _p(a), _p(b), a.shape[0], _p(out)))
        return out
