"""Multi-GPU layer of the C ABI (klb_multi_*: one process, several devices; klb_gather_*: one process per GPU, CUDA
IPC): chains sharded in contiguous blocks, RNG keyed by the global chain index, closing all-gather by the copy engines.
run(job::Vector{MCJob}) = map(run, job), src/jobs/jobs.jl:212.  1 device vs N devices must be bit-identical; on a
one-GPU box the same ordinal is listed several times, which exercises everything but the NVLink hop."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from helpers import assert_same, build_pair, synthetic_x0

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _devices(K, n):
    have = K._lib.lib().klb_device_count()
    return [g % have for g in range(n)]


@pytest.mark.parametrize("sampler,target,dim,nshards", [("HMC", "iso", 1024, 3), ("MALA", "rosen", 96, 2), ("MH", "iso", 7, 4),
                                                        ("HMC", "dense", 128, 2), ("HMC", "logit", 4, 5), ("NUTS", "shifted", 130, 3)])
def test_multi_equals_single_device(K, sampler, target, dim, nshards):
    L = K._lib
    N = 101                                               # ragged shards
    kw = dict(nchains=N, dim=dim, nsteps=24, burnin=7, thinning=2, step={"HMC": 0.02, "MALA": 0.002, "MH": 0.1, "NUTS": 0.1}[sampler],
              nleaps=5, seed=4242, sigma=np.full(dim, 0.05), tuner="dualavg" if sampler == "NUTS" else "accrate", period=5,
              target_rate=0.7, nadapt=15, maxndoublings=3)
    one, cfg, x0, tp, sg = build_pair(K, sampler, target, **kw)
    one.run()
    ref = one.output()
    tgt = one.parameter.target
    smp, tun, rng = one.sampler, one.tuner, one.range
    p = K.BasicContMuvParameter("p", logtarget=tgt)
    multi = K.BasicMCJob(K.likelihood_model(p, False), smp, rng, {"p": x0}, tuner=tun,
                         outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}, seed=4242,
                         devices=_devices(K, nshards))
    assert multi.ngpus == nshards
    multi.run()
    out = multi.output()
    assert_same("value", out.value, ref.value)
    assert_same("logtarget", out.logtarget, ref.logtarget)
    assert_same("accept", out.diagnosticvalues, ref.diagnosticvalues)
    assert_same("final state", multi.pstate_value, one.pstate_value)
    assert_same("tune.step", multi.tune.step, one.tune.step)
    for g in range(nshards):                              # every device holds the complete all-gather
        assert_same("gathered state on device %d" % g, multi.gathered(L.OUT_STATE, g), one.pstate_value)
        assert_same("gathered logtarget", multi.gathered(L.OUT_STATE_LOGTARGET, g), one.pstate_logtarget)
        assert_same("gathered step", multi.gathered(L.OUT_TUNE_STEP, g), one.tune.step)
        cnt = multi.gathered(L.OUT_TUNE_COUNTERS, g)
        assert_same("gathered accepted", cnt[:, 0], one.tune.accepted)
        assert_same("gathered totproposed", cnt[:, 2], one.tune.totproposed)
    assert_same("ess", multi.ess(), one.ess())
    assert_same("acceptance", multi.acceptance(), one.acceptance())
    if kw["tuner"] == "dualavg":
        # reset of a dual-averaging job that has run is refused on every shard, as for one device
        for jb in (multi, one):
            with pytest.raises(L.KlaraError) as ei:
                jb.reset()
            assert ei.value.code == L.KLB_EUNSUPPORTED
    else:
        # reset + second run continue the same streams on every shard
        multi.reset(); multi.run(); one.reset(); one.run()
        assert_same("second run", multi.output().value, one.output().value)
    multi.close()


def test_multi_real_devices_when_present(K):
    """two (or more) physical devices: the closing all-gather crosses NVLink; skipped on a one-GPU box"""
    L = K._lib
    have = L.lib().klb_device_count()
    if have < 2:
        pytest.skip("one device")
    N, d = 4096, 1024
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    mk = lambda **kw: K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 10), K.BasicMCRange(nsteps=20, burnin=10),  # noqa: E731
                                   {"p": K.SyntheticNormal(N, d)}, outopts={"monitor": ["logtarget"], "diagnostics": ["accept"]},
                                   seed=77, **kw)
    one, multi = mk(), mk(ngpus=0)
    assert multi.ngpus == have
    one.run(); multi.run()
    assert_same("logtarget", multi.output().logtarget, one.output().logtarget)
    for g in range(have):
        assert_same("gathered state on device %d" % g, multi.gathered(L.OUT_STATE, g), one.pstate_value)


def test_synthetic_state_and_seek(K, O):
    """klb_job_set_state_synthetic = the Philox initial value of SURVEY 8d (a function of the global chain index);
    klb_job_seek repositions the streams"""
    N, d = 37, 130
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    mk = lambda v0, **kw: K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 4), K.BasicMCRange(nsteps=9, burnin=2), {"p": v0},  # noqa: E731
                                       outopts={"monitor": ["value"]}, seed=99, **kw)
    a = mk(K.SyntheticNormal(N, d), chain_offset=1000)
    assert_same("x0", a.pstate_value, synthetic_x0(99, N, d, 1000))
    b = mk(synthetic_x0(99, N, d, 1000), chain_offset=1000)
    a.run(); b.run()
    assert_same("run from the synthetic state", a.output().value, b.output().value)
    first = a.output().value
    a.reset_synthetic(); a.seek(0); a.run()
    assert_same("seek(0) replays the run", a.output().value, first)
    a.reset_synthetic(); a.seek(500); a.run()
    assert not np.array_equal(a.output().value, first)
    cfg = O.make_config(O.HMC, O.ISO, N, d, 9, 2, step=0.05, nleaps=4, monitor=1, seed=99, chain_offset=1000, t0=500,
                        nv=a.plan().nv, nthreads=O.max_threads())
    assert_same("seek(500) == oracle at t0 = 500", a.output().value, O.run(cfg, synthetic_x0(99, N, d, 1000))["value"])
    odd = K.BasicMCJob(K.likelihood_model(p, False), K.MH(np.full(7, 0.3)), K.BasicMCRange(nsteps=3), {"p": K.SyntheticNormal(5, 7)}, seed=3)
    assert_same("odd dim", odd.pstate_value, synthetic_x0(3, 5, 7))


def test_device_peaks_are_plausible(K):
    fp64, dmma = K.device_peak("fp64"), K.device_peak("dmma")
    assert 5e12 < fp64 < 4e13, fp64                       # B200: 148 SMs x 64 lanes x ~1.9 GHz = 1.8e13 results/s
    assert 1e13 < dmma < 1.5e14, dmma


def _ipc_worker(rank, world, N, d, conn, ret):
    sys.path.insert(0, ROOT)
    import klara_b200 as K
    L = K._lib
    lib = L.lib()
    dev = rank % lib.klb_device_count()
    lo, hi = K.distributed.shard_range(N, rank, world)
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    job = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 5), K.BasicMCRange(nsteps=12, burnin=4),
                       {"p": K.SyntheticNormal(hi - lo, d)}, outopts={"monitor": ["logtarget"]}, seed=2024, device=dev,
                       chain_offset=lo)
    g = C.c_void_p()
    L.check(lib.klb_gather_create(job._h, world, rank, N, 0, C.byref(g)))
    h = C.create_string_buffer(L.GATHER_HANDLE_BYTES)
    L.check(lib.klb_gather_handle(g, h))
    conn.send(h.raw)                                      # "all-gather" of the handles through the parent
    allh = conn.recv()
    L.check(lib.klb_gather_connect(g, C.create_string_buffer(allh, len(allh))))
    job.run_async()
    L.check(lib.klb_gather_push_async(g))
    L.check(lib.klb_gather_sync(g))
    conn.send(b"pushed")                                  # inter-process barrier: every rank has pushed
    conn.recv()
    full = np.empty((N, d))
    L.check(lib.klb_gather_output(g, L.OUT_STATE, full.ctypes.data_as(C.c_void_p), full.nbytes))
    ret.put((rank, full))
    L.check(lib.klb_gather_disconnect(g))
    conn.recv()                                           # every rank has unmapped its peers before any buffer is freed
    lib.klb_gather_destroy(g)
    job.close()


def test_gather_between_processes_over_cuda_ipc(K):
    """one process per GPU (the torchrun layout of bench.py): IPC handles exchanged by the caller, every rank ends up
    with every chain's final state"""
    import multiprocessing as mp
    N, d, world = 64, 1024, 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    pipes = [ctx.Pipe() for _ in range(world)]
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, N, d, pipes[r][1], ret)) for r in range(world)]
    for pr in procs:
        pr.start()

    def recv(r):                                          # a worker that died must fail the test, not hang it
        assert pipes[r][0].poll(180), "rank %d did not answer (exit code %s)" % (r, procs[r].exitcode)
        return pipes[r][0].recv()
    try:
        handles = b"".join(recv(r) for r in range(world))
        for r in range(world):
            pipes[r][0].send(handles)
        for r in range(world):
            assert recv(r) == b"pushed"
        for r in range(world):
            pipes[r][0].send(b"go")
        got = dict(ret.get(timeout=120) for _ in range(world))
        for r in range(world):
            pipes[r][0].send(b"done")
    finally:
        for pr in procs:
            pr.join(timeout=60)
            if pr.is_alive():
                pr.kill()
    assert all(pr.exitcode == 0 for pr in procs)
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    one = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(0.05, 5), K.BasicMCRange(nsteps=12, burnin=4),
                       {"p": K.SyntheticNormal(N, d)}, outopts={"monitor": ["logtarget"]}, seed=2024)
    one.run()
    for r in range(world):
        assert_same("rank %d holds the whole all-gather" % r, got[r], one.pstate_value)


def test_gibbs_driver_matches_oracle(K, O):
    """BasicGibbsJob as a caller of the hot path (src/jobs/BasicGibbsJob.jl:185-231): every sweep runs each block's
    BasicMCJob (one launch), saves the dependent variables' states, resets the dpjobs.  Two blocks (HMC + tuned MALA)
    and a transformation of both, against oracle/gibbs.py bit for bit."""
    from oracle import gibbs as OG
    N = 19
    kw1 = dict(nchains=N, dim=130, nsteps=3, burnin=0, step=0.05, nleaps=4, seed=11, monitor=("value",), diagnostics=())
    kw2 = dict(nchains=N, dim=6, nsteps=5, burnin=5 - 1, step=0.3, seed=12, tuner="accrate", period=2, target_rate=0.6,
               monitor=("value",), diagnostics=())
    j1, c1, x1, tp1, sg1 = build_pair(K, "HMC", "iso", **kw1)
    j2, c2, x2, tp2, sg2 = build_pair(K, "MALA", "shifted", **kw2)
    a = K.BasicContMuvParameter("a", logtarget=j1.parameter.target)
    b = K.BasicContMuvParameter("b", logtarget=j2.parameter.target)
    tr = lambda v: np.concatenate([v["a"][:, :2] + v["b"][:, :2], v["a"].sum(1, keepdims=True)], axis=1)  # noqa: E731
    model = K.GenericModel([a, K.Transformation("s", tr), b], isindexed=False)
    job = K.BasicGibbsJob(model, {"a": j1, "b": j2}, K.BasicMCRange(nsteps=9, burnin=3, thinning=2), {"a": x1, "b": x2})
    job.run()
    got = job.output_dict()
    ref = OG.run_gibbs({"a": dict(cfg=c1, x0=x1, tparams=tp1), "b": dict(cfg=c2, x0=x2, tparams=tp2)}, {"s": tr},
                       ["a", "s", "b"], 9, 3, 2)
    assert got["a"].value.shape == (N, 3, 130) and got["s"].value.shape == (N, 3, 3)
    for k in ("a", "s", "b"):
        assert_same("gibbs " + k, got[k].value, ref[k])
    # the transformation saw the states of the SAME sweep: a was already updated, b not yet
    assert not np.array_equal(got["a"].value[:, 0], got["a"].value[:, 1])


def test_run_host_in_two_halves_and_over_devices(K):
    """klb_job_run_host_async / _finish (other calls are refused while a run is pending) and klb_multi_run_host (every
    device runs its shard's pipeline concurrently): same bits as set_state + run + output on one device"""
    L = K._lib
    lib = L.lib()
    N, d = 203, 130
    kw = dict(nchains=N, dim=d, nsteps=24, burnin=7, thinning=2, step=0.05, nleaps=5, seed=314, tuner="accrate", period=5, target_rate=0.7)
    ref_job, cfg, x0, tp, sg = build_pair(K, "HMC", "shifted", **kw)
    ref_job.run()
    ref = ref_job.output()
    job, *_ = build_pair(K, "HMC", "shifted", **kw)
    val, st = np.empty_like(ref.value), np.empty((N, d))
    arr = (L.KlbHostField * 2)()
    for i, (fld, buf) in enumerate(((L.OUT_VALUE, val), (L.OUT_STATE, st))):
        arr[i].field, arr[i].host_dst, arr[i].nbytes = fld, buf.ctypes.data, buf.nbytes
    L.check(lib.klb_job_run_host_async(job._h, x0.ctypes.data_as(C.c_void_p), arr, 2, 5))
    assert lib.klb_job_run(job._h) == L.KLB_ESTATE and lib.klb_job_reset(job._h) == L.KLB_ESTATE
    assert lib.klb_job_run_host_async(job._h, None, arr, 2, 5) == L.KLB_ESTATE
    L.check(lib.klb_job_run_host_finish(job._h))
    assert lib.klb_job_run_host_finish(job._h) == L.KLB_ESTATE
    assert_same("value", val, ref.value)
    assert_same("state", st, ref_job.pstate_value)
    # the logical job over four shards (ragged: 203 chains)
    p = K.BasicContMuvParameter("p", logtarget=ref_job.parameter.target)
    multi = K.BasicMCJob(K.likelihood_model(p, False), ref_job.sampler, ref_job.range, {"p": x0[::-1].copy()}, tuner=ref_job.tuner,
                         outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]}, seed=314, devices=_devices(K, 4))
    bufs = {L.OUT_VALUE: np.empty_like(ref.value), L.OUT_ACCEPT: np.empty_like(ref.diagnosticvalues), L.OUT_STATE: np.empty((N, d)),
            L.OUT_TUNE_STEP: np.empty(N)}
    multi.run_host(x0, bufs, 3)
    assert_same("multi value", bufs[L.OUT_VALUE], ref.value)
    assert_same("multi accept", bufs[L.OUT_ACCEPT], ref.diagnosticvalues)
    assert_same("multi state", bufs[L.OUT_STATE], ref_job.pstate_value)
    assert_same("multi step", bufs[L.OUT_TUNE_STEP], ref_job.tune.step)
    assert_same("gathered after run_host", multi.gathered(L.OUT_STATE, 3), ref_job.pstate_value)
    bad = x0.copy()
    bad[150, 3] = np.nan
    with pytest.raises(K.KlaraError) as ei:
        multi.run_host(bad, {}, 2)
    assert ei.value.code == L.KLB_ENOTFINITE and "chain 150" in str(ei.value)
    assert all(K._lib.lib().klb_job_plan(h, C.byref(pl)) == 0 and pl.transitions_done == 24
               for h, _, _ in multi._shards for pl in [L.KlbPlan()])       # no shard kept the rejected run
    multi.run_host(x0, bufs, 2)                          # the job is usable again: transitions 25..48 of the same streams
    ref_job.reset(x0)
    ref_job.run()
    assert_same("multi value after a rejected start", bufs[L.OUT_VALUE], ref_job.output().value)
