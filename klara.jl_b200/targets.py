"""Target descriptors: device-resident replacements of the `logtarget` / `gradlogtarget` closures a
user hands to Klara's BasicContMuvParameter (src/variables/parameters/BasicContMuvParameter.jl:383-411).

An arbitrary host closure cannot run inside a CUDA kernel, so the boundary carries a descriptor.
Each descriptor is also an ordinary callable on a vector (`t(z)`, `t.gradient(z)`), i.e. the same
object is a valid plain `logtarget` function for a host implementation.
"""
import numpy as np

from . import _lib as L


class Target:
    code = None

    def params(self, dim):
        """[(param id, float64 array)] uploaded through klb_job_set_target_f64."""
        return []


class IsoGaussian(Target):
    """plogtarget(z) = -dot(z, z);  pgradlogtarget(z) = -2*z        (README.md:153-155)"""
    code = L.TARGET_ISO

    def __call__(self, z):
        z = np.asarray(z, dtype=np.float64)
        return -float(np.dot(z, z))

    def gradient(self, z):
        return -2 * np.asarray(z, dtype=np.float64)


class ShiftedIsoGaussian(Target):
    """-(z-mu).(z-mu), gradient -2(z-mu)        (test/BasicContMuvParameter.jl:539-563)"""
    code = L.TARGET_SHIFTED_ISO

    def __init__(self, mu):
        self.mu = np.ascontiguousarray(mu, dtype=np.float64)

    def __call__(self, z):
        d = np.asarray(z, dtype=np.float64) - self.mu
        return -float(np.dot(d, d))

    def gradient(self, z):
        return -2 * (np.asarray(z, dtype=np.float64) - self.mu)

    def params(self, dim):
        if self.mu.size != dim:
            raise AssertionError("mu has %d entries, parameter has %d" % (self.mu.size, dim))
        return [(L.PARAM_MU, self.mu)]


class Rosenbrock(Target):
    """Paired Rosenbrock ("banana"), this repo's definition:
    logtarget = -scale * sum_k [ b (x[2k+1] - x[2k]^2)^2 + (a - x[2k])^2 ]   (0-based pairs)."""
    code = L.TARGET_ROSENBROCK

    def __init__(self, a=1.0, b=100.0, scale=0.05):
        self.a, self.b, self.scale = float(a), float(b), float(scale)

    def __call__(self, z):
        z = np.asarray(z, dtype=np.float64)
        x, y = z[0::2], z[1::2]
        return -self.scale * float(np.sum(self.b * (y - x * x) ** 2 + (self.a - x) ** 2))

    def gradient(self, z):
        z = np.asarray(z, dtype=np.float64)
        x, y = z[0::2], z[1::2]
        u = y - x * x
        g = np.empty_like(z)
        g[0::2] = self.scale * (4 * self.b * x * u + 2 * (self.a - x))
        g[1::2] = -self.scale * 2 * self.b * u
        return g

    def params(self, dim):
        return [(L.PARAM_ROSEN, np.array([self.a, self.b, self.scale], dtype=np.float64))]


class DenseGaussian(Target):
    """-z'Cz, gradient -2Cz with a dense precision matrix C: the closures of
    doc/examples/BivariateNormal/MALA/function/analytical.jl:4-21,

        logtarget     = (p, v) -> -dot(p, v[1]*p)
        gradlogtarget = (p, v) -> -2*v[1]*p            nkeys = 2,  model = GenericModel([C, p], isindexed=false)

    where `v` holds the values of ALL model vertices in vertex order (BasicContMuvParameter.jl:497-501: `nkeys` is
    only tested for > 0), so v[1] is the state of the `Hyperparameter(:C)` vertex.  `DenseGaussian(C)` carries the
    matrix itself; `DenseGaussian()` is bound by BasicMCJob from v0[:C] like the reference does."""
    code = L.TARGET_DENSE
    nkeys = 2

    def __init__(self, C=None):
        self.C = None
        if C is not None:
            self.bind([C])

    def bind(self, values):
        """values = [C]: the states of the model's other vertices, in vertex order"""
        if len(values) != 1:
            raise AssertionError("DenseGaussian reads one hyper-parameter (the precision matrix), the model supplies %d" % len(values))
        C = np.ascontiguousarray(values[0], dtype=np.float64)
        if C.ndim != 2 or C.shape[0] != C.shape[1]:
            raise AssertionError("the precision matrix must be square")
        self.C = C
        return self

    def __call__(self, z):
        z = np.asarray(z, dtype=np.float64)
        return -float(z @ (self.C @ z))

    def gradient(self, z):
        return -2 * (self.C @ np.asarray(z, dtype=np.float64))

    def params(self, dim):
        if self.C is None:
            raise AssertionError("DenseGaussian has no precision matrix: pass C or give v0 a value for the model's "
                                 "Hyperparameter vertex")
        if self.C.shape[0] != dim:
            raise AssertionError("C is %d x %d, parameter has %d entries" % (self.C.shape + (dim,)))
        return [(L.PARAM_C, self.C.reshape(-1))]


class BayesLogit(Target):
    """Bayesian logistic regression with a N(0, lambda I) prior on the coefficients: the closures of
    doc/examples/swiss/HMC/noadaptation/analytical.jl:11-20 (also swiss/MALA/analytical.jl), whose
    hyper-parameters arrive as `v = [lambda, X, y, p]` in model-vertex order
    (`likelihood_model([Hyperparameter(:λ), Data(:X), Data(:y), p], isindexed=false)`):

        ploglikelihood(p, v) = dot(Xp, y) - sum(log.(1+exp.(Xp))),   Xp = X*p
        plogprior(p, v)      = -0.5*(dot(p, p)/lambda + length(p)*log(2*pi*lambda))
        pgradlogtarget(p, v) = X'*(y - 1./(1+exp.(-X*p))) - p/lambda

    `BayesLogit(X, y, lam)` carries the data itself; `BayesLogit()` is bound by BasicMCJob from the
    initial values of the model's other vertices (v0[:λ], v0[:X], v0[:y]) like the reference does
    (BasicContMuvParameter.jl:497-501).  X is ndata x dim, dim <= 16."""
    code = L.TARGET_LOGIT
    nkeys = 4

    def __init__(self, X=None, y=None, lam=None):
        self.X = self.y = self.lam = None
        if X is not None:
            self.bind([lam, X, y])

    def bind(self, values):
        """values = [lambda, X, y]: the states of the other model vertices, in vertex order"""
        lam, X, y = values
        self.lam = 100.0 if lam is None else float(lam)
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        if self.X.ndim != 2 or self.X.shape[0] != self.y.size:
            raise AssertionError("X must be ndata x dim and y must have ndata entries")
        return self

    def loglikelihood(self, p):
        Xp = self.X @ np.asarray(p, dtype=np.float64)
        return float(np.dot(Xp, self.y) - np.sum(np.log(1 + np.exp(Xp))))

    def logprior(self, p):
        p = np.asarray(p, dtype=np.float64)
        return float(-0.5 * (np.dot(p, p) / self.lam + p.size * np.log(2 * np.pi * self.lam)))

    def __call__(self, p):
        return self.loglikelihood(p) + self.logprior(p)

    def gradient(self, p):
        p = np.asarray(p, dtype=np.float64)
        return self.X.T @ (self.y - 1. / (1 + np.exp(-(self.X @ p)))) - p / self.lam

    def params(self, dim):
        if self.X is None:
            raise AssertionError("BayesLogit has no data: pass X, y, lam or give v0 values for the model's "
                                 "Hyperparameter / Data vertices")
        if self.X.shape[1] != dim:
            raise AssertionError("X has %d columns, parameter has %d entries" % (self.X.shape[1], dim))
        return [(L.PARAM_LOGIT_LAMBDA, np.array([self.lam])), (L.PARAM_LOGIT_X, self.X.reshape(-1)),
                (L.PARAM_LOGIT_Y, self.y)]
