"""Multi-GPU host logic: one process per GPU, photon ids sharded contiguously, one small all-reduce.

The reference's only shared writes are its tallies (two `omp critical` adds,
/root/reference/src/polycap-source.c:963-980).  Here every rank traces a contiguous range of global
photon ids (same Philox seed, disjoint counters), and the per-energy weight sums (double[nE]) and the
int64 counter block are combined with one all-reduce each (NCCL over NVLink on GPUs; gloo in the CPU
tests).  Exit/leak records never cross ranks.
"""
import numpy as np


def shard_range(n_ids, rank, world_size, id0=0):
    """Contiguous id range of `rank`: [lo, hi).  Ranges tile [id0, id0+n_ids) exactly and differ by at most one id."""
    if n_ids < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad shard request")
    base, extra = divmod(n_ids, world_size)
    lo = id0 + rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def allreduce_tallies(sum_weights, counters, group=None):
    """In-place SUM all-reduce of the two tally tensors (torch tensors on the rank's device)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sum_weights, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return sum_weights, counters


def efficiencies(sum_weights, counters):
    """source.c:1066,1075: eff = sum_w/(exit+absorbed) * open_area_simulated = sum_w / started"""
    c = np.asarray(counters)
    started = float(c[1] + c[2] + c[3])  # exit + not_entered + not_transmitted
    return np.asarray(sum_weights, dtype=np.float64) / started if started > 0 else np.zeros(len(sum_weights))
