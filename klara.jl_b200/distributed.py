"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch on the
B200 box, gloo on CPU for the host-logic tests).

Chains are independent units (the reference's only multi-job facility is
`run(job::Vector) = map(run, job)`, src/jobs/jobs.jl:212), so they shard over ranks with no
data-path collective: rank r of R owns the contiguous block [r*N/R, (r+1)*N/R) of the
`dim x N` state matrix.  RNG streams are keyed by the GLOBAL chain index, so the sampled values do
not depend on R.  The only collective is one all-gather after the last transition.
"""
import numpy as np


def shard_range(nchains, rank, world):
    """[lo, hi) of the chains owned by `rank`; blocks differ by at most one chain."""
    if not (0 <= rank < world):
        raise ValueError("rank %d not in [0, %d)" % (rank, world))
    base, rem = divmod(nchains, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(nchains, world):
    return [shard_range(nchains, r, world)[1] - shard_range(nchains, r, world)[0] for r in range(world)]


def all_gather_chains(local, nchains, group=None):
    """all-gather per-chain rows (first axis = this rank's chains) into the full (nchains, ...) array.
    `local`: torch tensor (CUDA -> NCCL, CPU -> gloo) or numpy array (goes through a CPU tensor)."""
    import torch
    import torch.distributed as dist
    is_np = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local.contiguous()
    world = dist.get_world_size(group)
    sizes = shard_sizes(nchains, world)
    if t.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError("rank holds %d chains, expected %d" % (t.shape[0], sizes[dist.get_rank(group)]))
    if len(set(sizes)) == 1:
        out = torch.empty((nchains,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
    else:   # ragged shards: pad to the largest block
        m = max(sizes)
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
    return out.numpy() if is_np else out
