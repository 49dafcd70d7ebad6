"""The independent numpy / libm twin (oracle/twin.py) against the C oracle.  The twin shares no source with
klb_math.h / klb_oracle.c: libm exp/log, numpy (BLAS-order) dot products, Julia's evaluation order.  What these tests
pin: (i) the RNG contract, implemented twice from the published algorithms, gives the same bits; (ii) every sampler x
target of the C oracle agrees with the twin to north_star's tolerance -- accept/reject identical, log-target,
gradient and values within 1e-6 relative -- so the oracle's canonical reduction order, its klb_exp / klb_log and its
exact rewrites stand for the reference's unspecified BLAS order and libm."""
import math

import numpy as np
import pytest

from helpers import SAMPLERS, ar1_precision, logit_data
from oracle import twin as T
from twin_helpers import check_against_twin, relerr, twin_cfg, twin_target


def test_twin_philox_random123_kat():
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(v) for v in T.philox4x32(*ctr, *key, rounds=10)] == want
    assert T.ROUNDS == 7


def test_twin_ziggurat_table_is_the_published_construction():
    """256 layers of equal area v under exp(-x^2/2) (Marsaglia & Tsang 2000): x[i] (f(x[i+1]) - f(x[i])) = v,
    r = x[1] = 3.65415288536..., k[i] = floor(2^52 x[i+1] / x[i])"""
    x, k, f = T._ZX, T._ZK, T._ZF
    assert T._ZR == x[1] == pytest.approx(3.6541528853610088, rel=1e-15)
    fx = np.exp(-0.5 * x * x)
    assert relerr(f[1:256], fx[1:]) < 1e-15 and f[256] == 1.0
    xn = np.append(x[1:], 0.0)
    area = x * (np.append(fx[1:], 1.0) - np.append(0.0, fx[1:]))            # layer i spans f(x[i]) .. f(x[i+1]) (f(x[0]) := 0 for the base)
    v = x[1] * fx[1] + math.sqrt(math.pi / 2) * math.erfc(x[1] / math.sqrt(2))
    assert x[0] == pytest.approx(v / fx[1], rel=1e-14)
    assert relerr(area[1:], np.full(255, v)) < 1e-12
    kk = (k & np.uint64((1 << 52) - 1)).astype(np.float64)
    assert np.all(np.abs(kk - np.floor(2.0 ** 52 * xn / x)) <= 2)


@pytest.mark.parametrize("seed,chain,t", [(0, 0, 0), (20240925, 65535, 200), (2 ** 63 + 5, 2 ** 31 + 7, 2 ** 33 + 1)])
def test_twin_rng_contract_matches_oracle_bit_for_bit(O, seed, chain, t):
    """two implementations of the contract (numpy + libm here, klb_math.h there) give the same normals and uniforms"""
    n = 20001
    a, b = T.normals(seed, chain, t, n), O.normals(seed, chain, t, n)
    _same_normals(a, b)
    assert T.accept_uniform(seed, chain, t) == O.uniform(seed, chain, t)


def _same_normals(a, b):
    """bit-identical, except that a draw from the tail (|x| > r: x = r - log(u)/r) may differ in the last place,
    because the twin takes its logarithm from libm and the contract from klb_log (both < 1 ulp)"""
    diff = a.view(np.uint64) != b.view(np.uint64)
    assert np.all(np.abs(a[diff]) > T._ZR) and diff.sum() <= max(1, a.size // 50000)
    assert np.all(np.abs(a[diff] - b[diff]) <= 2 * np.spacing(np.abs(b[diff])))


def test_twin_slow_path_is_exercised(O):
    """enough draws to visit the wedges and the tail (|x| > r): both implementations still agree"""
    z = np.concatenate([T.normals(77, c, 3, 50000) for c in range(8)])
    zo = np.concatenate([O.normals(77, c, 3, 50000) for c in range(8)])
    _same_normals(z, zo)
    assert (np.abs(z) > T._ZR).sum() >= 20


def _oracle_run(O, sampler, target, nchains, dim, nsteps, rng_seed=0, **kw):
    rng = np.random.default_rng(rng_seed)
    tp, sigma = None, None
    tcode = {"iso": O.ISO, "shifted": O.SHIFTED, "dense": O.DENSE, "rosen": O.ROSEN, "logit": O.LOGIT}[target]
    if target == "shifted":
        tp = rng.normal(size=dim)
    elif target == "dense":
        tp = ar1_precision(dim).reshape(-1)
    elif target == "rosen":
        tp = np.array([1.0, 100.0, 0.05])
    elif target == "logit":
        tp = O.logit_params(*logit_data(dim, rng))
    if sampler == "MH":
        sigma = np.full(dim, kw.pop("sigma", 0.3))
    seed = kw.get("seed", 1234)
    x0 = np.stack([O.normals(seed, c, 0, dim) for c in range(nchains)]) * kw.pop("x0_scale", 1.0)
    tuner = kw.get("tuner", "vanilla")
    cfg = O.make_config(SAMPLERS[sampler], tcode, nchains, dim, nsteps, kw.get("burnin", 0), kw.get("thinning", 1),
                        kw.get("step", 0.1), kw.get("nleaps", 10),
                        {"vanilla": O.VANILLA, "accrate": O.ACCRATE, "dualavg": O.DUALAVG}[tuner],
                        kw.get("target_rate", 0.574), 7.0, kw.get("period", 100), int(kw.get("verbose", False)),
                        7 if sampler != "MH" else 3, 1, seed, 0, 0, 0, None, O.max_threads(), nadapt=kw.get("nadapt", 1000))
    ref = O.run(cfg, x0, tp, sigma)
    return ref, x0, tp, twin_cfg(sampler, nsteps, sigma=sigma, **kw)


CASES = [
    ("HMC", "iso", 1024, dict(step=0.05, nleaps=10)),
    ("HMC", "iso", 7, dict(step=0.2, nleaps=5)),
    ("HMC", "shifted", 100, dict(step=0.08, nleaps=6)),
    ("HMC", "dense", 64, dict(step=0.05, nleaps=8, x0_scale=0.3)),
    ("HMC", "rosen", 96, dict(step=0.01, nleaps=6)),
    ("HMC", "logit", 4, dict(step=0.02, nleaps=5)),
    ("MALA", "iso", 128, dict(step=0.05)),
    ("MALA", "shifted", 33, dict(step=0.1)),
    ("MALA", "dense", 32, dict(step=0.02, x0_scale=0.3)),
    ("MALA", "rosen", 256, dict(step=0.002)),
    ("MALA", "logit", 4, dict(step=0.005)),
    ("MH", "iso", 2, dict(sigma=1.0)),
    ("MH", "iso", 1024, dict(sigma=0.02)),
    ("MH", "shifted", 65, dict(sigma=0.1)),
    ("MH", "dense", 16, dict(sigma=0.1, x0_scale=0.3)),
    ("MH", "rosen", 32, dict(sigma=0.05)),
    ("MH", "logit", 5, dict(sigma=0.05)),
]


@pytest.mark.parametrize("sampler,target,dim,kw", CASES, ids=["%s-%s-%d" % c[:3] for c in CASES])
def test_oracle_agrees_with_twin_free_running(O, sampler, target, dim, kw):
    """whole chains, both sides free-running: identical accept/reject sequences, values / log-targets / gradients
    within 1e-6 relative (observed: ~1e-13)"""
    nchains, nsteps = 4, 40
    ref, x0, tp, cfg = _oracle_run(O, sampler, target, nchains, dim, nsteps, **dict(kw))
    w = check_against_twin("oracle", cfg, twin_target(target, dim, tp), x0, range(nchains), ref["value"], ref["logtarget"],
                           ref["accept"], grad=ref["gradlogtarget"] if sampler != "MH" else None)
    assert w["flips"] == 0 and w["value"] < 1e-9 and w["logtarget"] < 1e-9
    assert 0 < ref["accept"].mean() <= 1.0


@pytest.mark.parametrize("sampler,step", [("HMC", 0.07), ("MALA", 0.3)])
def test_oracle_tuners_agree_with_twin(O, sampler, step):
    """AcceptanceRateMCTuner: the per-chain step after burn-in follows the same sequence of tune! events"""
    nchains, dim = 3, 20
    kw = dict(step=step, nleaps=4, tuner="accrate", target_rate=0.7, period=10, burnin=60, seed=99)
    ref, x0, tp, cfg = _oracle_run(O, sampler, "iso", nchains, dim, 90, **kw)
    w = check_against_twin("oracle", cfg, T.IsoGaussian(), x0, range(nchains), ref["value"], ref["logtarget"],
                           ref["accept"], final_step=ref["tune"]["step"])
    assert w["flips"] == 0
    assert (ref["tune"]["step"] != step).all()


def test_oracle_dual_averaging_agrees_with_twin(O):
    nchains, dim = 3, 16
    kw = dict(step=0.1, nleaps=8, tuner="dualavg", target_rate=0.65, nadapt=30, seed=5)
    ref, x0, tp, cfg = _oracle_run(O, "HMC", "iso", nchains, dim, 50, **kw)
    w = check_against_twin("oracle", cfg, T.IsoGaussian(), x0, range(nchains), ref["value"], ref["logtarget"],
                           ref["accept"], final_step=ref["tune"]["step"])
    assert w["flips"] == 0
    assert len(np.unique(ref["da"]["nleaps"])) >= 1 and (ref["da"]["count"] == 50).all()


@pytest.mark.parametrize("name,sampler,target,dim,nsteps,kw", [
    ("C2", "MALA", "iso", 128, 300, dict(step=0.9)),
    ("C3", "HMC", "iso", 1024, 60, dict(step=0.05, nleaps=10)),
    ("C4", "HMC", "dense", 512, 6, dict(step=0.02, nleaps=20, x0_scale=0.3)),
    ("C5", "MALA", "rosen", 256, 300, dict(step=0.01, tuner="accrate", target_rate=0.574, period=100)),
])
def test_oracle_agrees_with_twin_teacher_forced_at_baseline_parameters(O, name, sampler, target, dim, nsteps, kw):
    """BASELINE.json's configurations at their stated sampler parameters, every transition restarted from the oracle's
    previous state so that a long (possibly chaotic) chain cannot compound rounding differences"""
    nchains = 2
    ref, x0, tp, cfg = _oracle_run(O, sampler, target, nchains, dim, nsteps, **dict(kw))
    w = check_against_twin(name, cfg, twin_target(target, dim, tp), x0, range(nchains), ref["value"], ref["logtarget"],
                           ref["accept"], forced=True)
    assert w["transitions"] == nchains * nsteps and w["flips"] == 0
