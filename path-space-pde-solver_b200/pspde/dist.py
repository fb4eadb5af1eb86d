"""Trajectory sharding over ranks (one process per GPU) and the two small collectives of a training iteration.

The reference is single-device (solver.py:36).  Trajectories are independent given theta, so the K paths are split
into contiguous blocks of the GLOBAL path index; the Philox counter is keyed on that global index, which makes
every per-path result independent of the number of ranks.  Collectives (NCCL on GPUs, gloo in the CPU tests):
  * loss statistics   all_reduce(sum) of a handful of fp64 scalars after the forward rollout
  * gradient          all_reduce(sum) of the flat fp32 gradient
"""
import torch as pt
import torch.distributed as td


def world(group=None):
    if td.is_available() and td.is_initialized():
        return td.get_rank(group), td.get_world_size(group)
    return 0, 1


def shard_range(K, rank, world_size):
    """Contiguous block [lo, hi) of the global path index for `rank`; the first K % W ranks get one extra path."""
    base, extra = divmod(int(K), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_sum_(t, group=None):
    if td.is_available() and td.is_initialized() and td.get_world_size(group) > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM, group=group)
    return t
