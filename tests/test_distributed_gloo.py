"""world_size-2 gloo test of the N > 1 host path: chains sharded over ranks by shard_range, each rank
advancing its shard with the global chain offset, one all-gather at the end; the gathered result must be
bit-identical to the single-process run.  (On CPU the per-rank compute is the oracle -- the CUDA library has no
CPU path -- so this covers exactly the host-side logic bench.py and the GPU ranks share.)"""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nchains, dim, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import klara_b200 as K
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = K.distributed.shard_range(nchains, rank, world)
    x0 = np.stack([O.normals(5, c, 0, dim) for c in range(lo, hi)])
    cfg = O.make_config(O.HMC, O.ISO, hi - lo, dim, 25, burnin=5, step=0.1, nleaps=3, monitor=1, seed=5,
                        chain_offset=lo)
    r = O.run(cfg, x0)
    full_state = K.distributed.all_gather_chains(r["x"], nchains)
    full_value = K.distributed.all_gather_chains(r["value"], nchains)
    if rank == 0:
        q.put((full_state, full_value))
    dist.barrier()
    dist.destroy_process_group()


def _run(nchains, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, world, port, nchains, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def _single(nchains):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    x0 = np.stack([O.normals(5, c, 0, 6) for c in range(nchains)])
    cfg = O.make_config(O.HMC, O.ISO, nchains, 6, 25, burnin=5, step=0.1, nleaps=3, monitor=1, seed=5)
    return O.run(cfg, x0)


def test_two_ranks_equal_one_process_even_split():
    state, value = _run(8)
    ref = _single(8)
    assert np.array_equal(state, ref["x"]) and np.array_equal(value, ref["value"])


def test_two_ranks_ragged_split():
    state, value = _run(7)
    ref = _single(7)
    assert np.array_equal(state, ref["x"]) and np.array_equal(value, ref["value"])
