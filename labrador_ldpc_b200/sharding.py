"""Multi-GPU partitioning of a batch of independent codewords.

Every codeword is independent (the reference zeroes its scratch on every call,
src/decoder.rs:368,374), so a batch is split into contiguous frame ranges, one
per rank / GPU, and there is NO collective on the data path.  The only
cross-rank traffic is the optional reduction of a handful of counters
(frames, successes, iteration histogram), mirroring the single AtomicU64 that
the reference's perftest uses to merge its workers (perftest/src/main.rs:37-56).
"""
import numpy as np


def shard_bounds(batch, world_size, rank):
    """Contiguous frame range [lo, hi) of `rank`; ranges tile [0, batch) exactly.

    Same rule as the C library's multi-device split (runtime.cu: batch*g/G).
    """
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    lo = batch * rank // world_size
    hi = batch * (rank + 1) // world_size
    return lo, hi


def shard_view(array, world_size, rank):
    """The rank's slice of a frame-major [batch, ...] array (no copy)."""
    lo, hi = shard_bounds(len(array), world_size, rank)
    return array[lo:hi]


class DecodeStats:
    """Counters that are summed across ranks: frames, successes, iteration histogram."""
    HIST = 128

    def __init__(self):
        self.frames = 0
        self.successes = 0
        self.iter_sum = 0
        self.hist = np.zeros(self.HIST, np.int64)

    def add(self, success, iters):
        success = np.asarray(success).astype(bool)
        iters = np.asarray(iters).astype(np.int64)
        self.frames += int(success.size)
        self.successes += int(success.sum())
        self.iter_sum += int(iters.sum())
        self.hist += np.bincount(np.minimum(iters, self.HIST - 1), minlength=self.HIST)
        return self

    def to_vector(self):
        return np.concatenate([[self.frames, self.successes, self.iter_sum], self.hist]).astype(np.int64)

    @classmethod
    def from_vector(cls, v):
        s = cls()
        s.frames, s.successes, s.iter_sum = int(v[0]), int(v[1]), int(v[2])
        s.hist = np.asarray(v[3:], np.int64).copy()
        return s

    def all_reduce(self, dist, device="cpu"):
        """Sum over ranks with torch.distributed (gloo on CPU, nccl on GPU)."""
        import torch
        t = torch.from_numpy(self.to_vector()).to(device)
        dist.all_reduce(t)
        return DecodeStats.from_vector(t.cpu().numpy())

    @property
    def fer(self):
        return 1.0 - self.successes / max(self.frames, 1)

    @property
    def mean_iters(self):
        return self.iter_sum / max(self.frames, 1)
