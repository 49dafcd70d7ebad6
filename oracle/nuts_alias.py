"""Two Python models of the reference's multivariate NUTS transition -- TEST INFRASTRUCTURE, never imported by the product.

`alias_transition` mirrors /root/reference/src/samplers/iterate/NUTS.jl:230-457 and build_tree!
(/root/reference/src/samplers/NUTS.jl:514-628 Vanilla, :781-927 DualAveraging) statement by statement, on Python objects
that have Julia's reference semantics: a ParameterState is a mutable object, `a.x = b.y` binds a reference, `v[:] = ...`
writes in place, a destructuring assignment evaluates its right-hand side first and then assigns left to right.  That
matters because the reference builds its sampler state as

    MuvNUTSState(pstate, tune) = MuvNUTSState(pstate, pstate, pstate, pstate, tune, ...)        NUTS.jl:198-225

i.e. pstateplus, pstateminus, pstateprime and pstatedprime are ONE object from the start, every leaf of build_tree! returns
sstate.pstateprime / sstate.momentumprime three times over (:533-539), and every level keeps its results in the one shared
sstate (:541-549).  What the reference computes is therefore not the textbook tree: it is decided by this aliasing.

`simple_transition` is the state machine that falls out when the aliasing is resolved by hand:
one moving point, one running momentum, a pristine copy of the initial momentum per direction until that direction is
first used.  tests/test_oracle_nuts.py checks that the two agree exactly (same draws, same floating-point operations) over
randomised configurations, and then pins oracle/klb_oracle.c's restatement of the state machine to it.

Random numbers come from a `draws` object with randn(d), rand() and randbool(); the models only fix the ORDER in which
the reference consumes them (randn for the momentum, rand for the slice variable, then per doubling rand(Bool), the rand()
of every inner node whose first half did not stop, and the rand() of the acceptance test when the subtree did not stop).
"""
import math

import numpy as np


class PState:
    def __init__(self, d):
        self.value = np.full(d, np.nan)
        self.gradlogtarget = np.full(d, np.nan)
        self.logtarget = math.nan


class Tune:
    def __init__(self, step):
        self.step = step
        self.accepted = 0
        self.proposed = 0


class SState:
    """MuvNUTSState(pstate, tune): the four states are the same object (NUTS.jl:198-225)"""

    def __init__(self, d, tune):
        e = PState(d)                                   # generate_empty(pstate, ...)
        self.pstateplus = self.pstateminus = self.pstateprime = self.pstatedprime = e
        self.tune = tune
        self.momentum = np.full(d, np.nan)
        self.momentumplus = np.full(d, np.nan)
        self.momentumminus = np.full(d, np.nan)
        self.momentumprime = np.full(d, np.nan)
        self.oldhamiltonian = math.nan
        self.u = math.nan
        self.v = 0
        self.j = 0
        self.n = 0
        self.nprime = 0
        self.ndprime = 0
        self.s = True
        self.sprime = True
        self.sdprime = True
        self.update = True
        self.count = 0


class Target:
    """-z.z with gradient -2z (README.md:153-155), or the shifted version; enough to exercise the control flow"""

    def __init__(self, mu=None):
        self.mu = mu

    def logtarget(self, ps):
        z = ps.value if self.mu is None else ps.value - self.mu
        ps.logtarget = -float(np.dot(z, z))

    def gradlogtarget(self, ps):
        z = ps.value if self.mu is None else ps.value - self.mu
        ps.gradlogtarget = -2.0 * z


class LogitTarget:
    """the closures of doc/examples/swiss/NUTS/*/analytical.jl:11-20, v = [lambda, X, y, p], in numpy / libm:
    ploglikelihood = dot(Xp, y) - sum(log.(1+exp.(Xp))), plogprior = -0.5*(dot(p, p)/lambda + length(p)*log(2*pi*lambda)),
    pgradlogtarget = X'*(y - 1./(1+exp.(-X*p))) - p/lambda;  logtarget = loglikelihood + logprior"""

    def __init__(self, X, y, lam):
        self.X, self.y, self.lam = np.asarray(X, dtype=np.float64), np.asarray(y, dtype=np.float64), float(lam)

    def logtarget(self, ps):
        p = ps.value
        Xp = self.X @ p
        loglik = float(np.dot(Xp, self.y)) - float(np.sum(np.log(1 + np.exp(Xp))))
        logprior = -0.5 * (float(np.dot(p, p)) / self.lam + p.size * np.log(2 * np.pi * self.lam))
        ps.logtarget = loglik + logprior

    def gradlogtarget(self, ps):
        p = ps.value
        ps.gradlogtarget = self.X.T @ (self.y - 1.0 / (1 + np.exp(-(self.X @ p)))) - p / self.lam


def hamiltonian(logtarget, momentum):                    # samplers.jl:103
    return logtarget - 0.5 * float(np.dot(momentum, momentum))


def leapfrog(pstate, momentum, pstate0, momentum0, step, target):   # samplers.jl:122-134
    momentum[:] = momentum0 + 0.5 * step * pstate0.gradlogtarget
    pstate.value[:] = pstate0.value + step * momentum
    target.gradlogtarget(pstate)
    momentum[:] = momentum + 0.5 * step * pstate.gradlogtarget


def uturn(xp, xm, mp, mm):                                # NUTS.jl:395-396
    return float(np.dot(xp - xm, mp)) < 0.0 or float(np.dot(xp - xm, mm)) < 0.0


def _min1exp(x):
    e = math.exp(x) if x < 700 else math.inf
    return e if e != e else min(1.0, e)


def build_tree(ss, pstate, momentum, oldh, u, v, j, target, maxdelta, draws, da):
    """NUTS.jl:514-628 (Vanilla) / :781-927 (DualAveraging: two more return values)"""
    if j == 0:
        leapfrog(ss.pstateprime, ss.momentumprime, pstate, momentum, v * ss.tune.step, target)
        target.logtarget(ss.pstateprime)
        hprime = hamiltonian(ss.pstateprime.logtarget, ss.momentumprime)
        ss.nprime = int(u <= hprime)
        ss.sprime = u < maxdelta + hprime
        return (ss.pstateprime, ss.momentumprime, ss.pstateprime, ss.momentumprime, ss.pstateprime, ss.nprime, ss.sprime,
                _min1exp(hprime - oldh) if da else None, 1)
    (ss.pstateminus, ss.momentumminus, ss.pstateplus, ss.momentumplus, ss.pstateprime, ss.nprime, ss.sprime, aprime, naprime) = \
        build_tree(ss, pstate, momentum, oldh, u, v, j - 1, target, maxdelta, draws, da)
    if ss.sprime:
        if v == -1:
            (ss.pstateminus, ss.momentumminus, _, _, ss.pstatedprime, ss.ndprime, ss.sdprime, adprime, nadprime) = \
                build_tree(ss, ss.pstateminus, ss.momentumminus, oldh, u, v, j - 1, target, maxdelta, draws, da)
        else:
            (_, _, ss.pstateplus, ss.momentumplus, ss.pstatedprime, ss.ndprime, ss.sdprime, adprime, nadprime) = \
                build_tree(ss, ss.pstateplus, ss.momentumplus, oldh, u, v, j - 1, target, maxdelta, draws, da)
        r = draws.rand()
        den = ss.ndprime + ss.nprime
        if den != 0 and r <= ss.ndprime / den:            # 0/0 is NaN in Julia: the comparison is false
            ss.pstateprime.value = ss.pstatedprime.value.copy()
            ss.pstateprime.gradlogtarget = ss.pstatedprime.gradlogtarget.copy()
            ss.pstateprime.logtarget = ss.pstatedprime.logtarget
        ss.nprime += ss.ndprime
        ss.sprime = ss.sdprime and not uturn(ss.pstateplus.value, ss.pstateminus.value, ss.momentumplus, ss.momentumminus)
        if da:
            aprime += adprime
        naprime += nadprime
    return (ss.pstateminus, ss.momentumminus, ss.pstateplus, ss.momentumplus, ss.pstateprime, ss.nprime, ss.sprime, aprime, naprime)


def alias_transition(pstate, ss, target, maxdelta, maxndoublings, draws, da=False):
    """iterate!(job, NUTS, Multivariate), iterate/NUTS.jl:230-400 (the tuner blocks that follow are the caller's).
    Returns (update, ndoublings, a, na)."""
    a = na = None
    ss.momentum[:] = draws.randn(pstate.value.size)
    ss.oldhamiltonian = hamiltonian(pstate.logtarget, ss.momentum)
    ss.pstateplus.value = pstate.value.copy()
    ss.pstateplus.gradlogtarget = pstate.gradlogtarget.copy()
    ss.momentumplus = ss.momentum.copy()
    ss.pstateminus.value = pstate.value.copy()
    ss.pstateminus.gradlogtarget = pstate.gradlogtarget.copy()
    ss.momentumminus = ss.momentum.copy()
    ss.j = 0
    ss.n = 1
    ss.s = True
    ss.update = False
    ss.u = math.log(draws.rand()) + ss.oldhamiltonian
    while ss.s and ss.j < maxndoublings:
        ss.v = 1 if draws.randbool() else -1
        if ss.v == -1:
            (ss.pstateminus, ss.momentumminus, _, _, ss.pstateprime, ss.nprime, ss.sprime, a, na) = \
                build_tree(ss, ss.pstateminus, ss.momentumminus, ss.oldhamiltonian, ss.u, ss.v, ss.j, target, maxdelta, draws, da)
        else:
            (_, _, ss.pstateplus, ss.momentumplus, ss.pstateprime, ss.nprime, ss.sprime, a, na) = \
                build_tree(ss, ss.pstateplus, ss.momentumplus, ss.oldhamiltonian, ss.u, ss.v, ss.j, target, maxdelta, draws, da)
        if ss.sprime and draws.rand() < ss.nprime / ss.n:
            pstate.value = ss.pstateprime.value.copy()
            pstate.gradlogtarget = ss.pstateprime.gradlogtarget.copy()
            pstate.logtarget = ss.pstateprime.logtarget
            ss.update = True
        ss.j += 1
        ss.n += ss.nprime
        ss.s = ss.sprime and not uturn(ss.pstateplus.value, ss.pstateminus.value, ss.momentumplus, ss.momentumminus)
    return ss.update, ss.j, a, na


# ------------------------------------------------------------------ the resolved state machine
def simple_tree(x, g, m, step_v, j, u, oldh, target, maxdelta, draws, da):
    """T(j) on the one moving point (x, g, running momentum m): up to 2^j leapfrog steps.  Returns (lt, n, s, a, na).
    After leaf l (0-based) the recursion unwinds through the levels k = 1..j: where the finished block was a second half
    (bit k-1 of l set) a rand() is drawn, n doubles, and the sums of the first half are added in front; where it was a first
    half the level goes on to its second half if s holds and returns as it is otherwise."""
    saved_a = [0.0] * (j + 1)
    saved_na = [0] * (j + 1)
    leaf = 0
    lt = math.nan
    while True:
        # leaf: NUTS.jl:527-539
        m[:] = m + 0.5 * step_v * g
        x[:] = x + step_v * m
        ps = PState(x.size)
        ps.value = x
        target.gradlogtarget(ps)
        g[:] = ps.gradlogtarget
        m[:] = m + 0.5 * step_v * g
        target.logtarget(ps)
        lt = ps.logtarget
        hprime = hamiltonian(lt, m)
        n = int(u <= hprime)
        s = u < maxdelta + hprime
        a = _min1exp(hprime - oldh) if da else None
        na = 1
        k = 1
        descend = False
        while k <= j:
            if (leaf >> (k - 1)) & 1:                    # a second half is complete
                draws.rand()
                n = 2 * n
                if da:
                    a = saved_a[k] + a
                na = saved_na[k] + na
                k += 1
            elif s:                                       # a first half is complete and did not stop: do the second
                saved_a[k] = a
                saved_na[k] = na
                descend = True
                break
            else:                                         # a first half stopped: the level returns it as it is
                k += 1
        if not descend:
            return lt, n, s, a, na
        leaf += 1


def simple_transition(state, step, target, maxdelta, maxndoublings, draws, da=False):
    """state: dict(value, gradlogtarget, logtarget) = job.pstate, updated in place.  Returns (update, ndoublings, a, na)."""
    d = state["value"].size
    p0 = draws.randn(d).copy()
    oldh = hamiltonian(state["logtarget"], p0)
    x = state["value"].copy()
    g = state["gradlogtarget"].copy()
    m = np.full(d, np.nan)                                # momentumprime: whatever the last transition left in it
    aliased = {1: False, -1: False}
    j, n, s, update = 0, 1, True, False
    u = math.log(draws.rand()) + oldh
    a = na = None
    while s and j < maxndoublings:
        v = 1 if draws.randbool() else -1
        if not aliased[v]:
            m[:] = p0                                     # the first leaf reads this direction's own copy of the momentum
        lt, nprime, sprime, a, na = simple_tree(x, g, m, v * step, j, u, oldh, target, maxdelta, draws, da)
        aliased[v] = True
        if j >= 1:
            aliased[-v] = True                            # NUTS.jl:541-549 rebinds both ends to momentumprime
        if sprime and draws.rand() < nprime / n:
            state["value"] = x.copy()
            state["gradlogtarget"] = g.copy()
            state["logtarget"] = lt
            update = True
        j += 1
        n += nprime
        s = sprime                                        # uturn(x - x, ...) = (0 < 0) is false
    return update, j, a, na
