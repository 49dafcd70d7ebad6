"""Shared helpers: run one configuration through the CUDA path (C ABI via the Python mirror) and
through the CPU oracle on the same seeded inputs, and compare bit for bit."""
import numpy as np

from oracle import oracle as O

SAMPLERS = {"MH": O.MH, "MALA": O.MALA, "HMC": O.HMC, "NUTS": O.NUTS}


def synthetic_x0(seed, nchains, dim, chain_offset=0):
    """x0[c, i] = N(0,1) from Philox stream (seed, chain, transition 0, element i)   """
    return np.stack([O.normals(seed, chain_offset + c, 0, dim) for c in range(nchains)])


def ar1_precision(dim, rho=0.8):
    """C = inv(Sigma), Sigma_ij = rho^|i-j|: the d-dim analogue of the reference's bivariate example
    (doc/examples/BivariateNormal/MALA/function/analytical.jl:21); made exactly symmetric"""
    idx = np.arange(dim)
    C = np.linalg.inv(rho ** np.abs(idx[:, None] - idx[None, :]))
    return np.ascontiguousarray((C + C.T) / 2)


def make_target(K, name, dim, rng):
    if name == "iso":
        return K.IsoGaussian(), O.ISO, None
    if name == "shifted":
        mu = rng.normal(size=dim)
        return K.ShiftedIsoGaussian(mu), O.SHIFTED, mu
    if name == "rosen":
        return K.Rosenbrock(1.0, 100.0, 0.05), O.ROSEN, np.array([1.0, 100.0, 0.05])
    if name == "dense":
        C = ar1_precision(dim)
        return K.DenseGaussian(C), O.DENSE, C.reshape(-1)
    if name == "logit":
        X, y, lam = logit_data(dim, rng)
        return K.BayesLogit(X, y, lam), O.LOGIT, O.logit_params(X, y, lam)
    raise KeyError(name)


def logit_data(dim, rng, ndata=200, lam=100.0):
    """Synthetic stand-in for the swiss bank-note data of doc/examples/swiss (200 x 4 standardised covariates,
    0/1 outcome): standardised Gaussian covariates, outcomes drawn from a logistic model"""
    X = rng.normal(size=(ndata, dim))
    X = (X - X.mean(0)) / X.std(0, ddof=1)
    beta = rng.normal(size=dim)
    y = (rng.uniform(size=ndata) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    return np.ascontiguousarray(X), y, lam


def build_pair(K, sampler, target, nchains, dim, nsteps, burnin=0, thinning=1, step=0.1, nleaps=10,
               tuner="vanilla", target_rate=0.574, period=100, verbose=False, monitor=("value", "logtarget"),
               diagnostics=("accept",), seed=1234, arith="reference", chain_offset=0, x0=None, sigma=None,
               device=0, rng_seed=0, nadapt=1000, score="logistic", maxdelta=1000, maxndoublings=5):
    rng = np.random.default_rng(rng_seed)
    tgt, tcode, tparams = make_target(K, target, dim, rng)
    if x0 is None:
        x0 = synthetic_x0(seed, nchains, dim, chain_offset)
    if sampler == "MH":
        sigma = np.full(dim, 0.5) if sigma is None else np.asarray(sigma, dtype=np.float64)
        smp = K.MH(sigma)
    elif sampler == "MALA":
        smp = K.MALA(step)
    elif sampler == "NUTS":
        smp = K.NUTS(step, maxdelta=maxdelta, maxndoublings=maxndoublings)
    else:
        smp = K.HMC(step, nleaps)
    da_kw = dict(nadapt=nadapt, eps0bar=1.0, h0bar=0.0, gamma=0.05, t0=10, kappa=0.75)
    tun = K.VanillaMCTuner(period=period, verbose=verbose) if tuner == "vanilla" else \
        K.DualAveragingMCTuner(target_rate, period=period, verbose=verbose, **da_kw) if tuner == "dualavg" else \
        K.AcceptanceRateMCTuner(target_rate, period=period, verbose=verbose,
                                score=K.erf_rate_score if score == "erf" else K.logistic_rate_score)
    p = K.BasicContMuvParameter("p", logtarget=tgt)
    model = K.likelihood_model(p, False)
    rng_ = K.BasicMCRange(nsteps=nsteps, burnin=burnin, thinning=thinning)
    outopts = {"monitor": list(monitor), "diagnostics": list(diagnostics)}
    job = K.BasicMCJob(model, smp, rng_, {"p": x0}, tuner=tun, outopts=outopts, seed=seed, arith=arith,
                       chain_offset=chain_offset, device=device)
    mon = sum({"value": 1, "logtarget": 2, "gradlogtarget": 4}[m] for m in monitor)
    cfg = O.make_config(SAMPLERS[sampler], tcode, nchains, dim, nsteps, burnin, thinning, step, nleaps,
                        {"vanilla": O.VANILLA, "accrate": O.ACCRATE, "dualavg": O.DUALAVG}[tuner], target_rate,
                        3.0 if score == "erf" else 7.0,
                        period, int(verbose), mon, sum({"accept": 1, "ndoublings": 2, "a": 4, "na": 8}[k] for k in set(diagnostics)),
                        seed, chain_offset, 0,
                        1 if arith == "fma" else 0, job.plan().nv, O.max_threads(), nadapt=nadapt, eps0bar=1.0, h0bar=0.0,
                        gamma=0.05, da_t0=10, kappa=0.75, score=1 if score == "erf" else 0, maxdelta=maxdelta,
                        maxndoublings=maxndoublings)
    return job, cfg, x0, tparams, sigma


def assert_same(name, a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (name, a.shape, b.shape)
    if a.dtype.kind == "f":
        same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    if not same.all():
        idx = np.argwhere(~same)[0]
        raise AssertionError("%s differs at %s: gpu=%r oracle=%r (%d of %d entries differ)"
                             % (name, tuple(idx), a[tuple(idx)], b[tuple(idx)], (~same).sum(), same.size))


def compare_run(job, cfg, x0, tparams, sigma, t0=0):
    """run both sides; every monitored field, the final state and the tuner records must be identical"""
    cfg.t0 = t0
    job.run()
    ref = O.run(cfg, x0, tparams, sigma)
    out = job.output()
    if cfg.monitor & 1:
        assert_same("value", out.value, ref["value"])
    if cfg.monitor & 2:
        assert_same("logtarget", out.logtarget, ref["logtarget"])
    if cfg.monitor & 4:
        assert_same("gradlogtarget", out.gradlogtarget, ref["gradlogtarget"])
    keys = list(out.diagnostickeys)
    if len(keys) == 1:
        assert_same(keys[0], out.diagnosticvalues, ref[keys[0]])
    for q, key in enumerate(keys if len(keys) > 1 else []):   # (nchains, nkeys, npost) in the order of outopts[:diagnostics]
        got = out.diagnosticvalues[..., q, :]
        assert_same(key, got, ref[key].astype(got.dtype))      # the stack is float64 when :a is among the keys
    assert_same("final state", job.pstate_value, ref["x"])
    assert_same("final logtarget", job.pstate_logtarget, ref["logtarget_state"])
    tn = job.tune
    assert_same("tune.step", tn.step, ref["tune"]["step"])
    assert_same("tune.accepted", tn.accepted, ref["tune"]["accepted"])
    assert_same("tune.proposed", tn.proposed, ref["tune"]["proposed"])
    assert_same("tune.totproposed", tn.totproposed, ref["tune"]["totproposed"])
    assert_same("tune.rate", tn.rate, ref["tune"]["rate"])
    if ref.get("da") is not None:
        for mine, theirs in (("lam", "lambda"), ("mu", "mu"), ("epsbar", "epsbar"), ("hbar", "hbar"), ("hweight", "hweight"),
                             ("epsweight", "epsweight")):
            assert_same("tune." + mine, getattr(tn, mine), ref["da"][theirs])
        assert_same("tune.nleaps", tn.nleaps, ref["da"]["nleaps"].astype(np.int64))
        assert_same("sstate.count", tn.count, ref["da"]["count"].astype(np.int64))
    return out, ref
