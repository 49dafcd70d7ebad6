"""BasicGibbsJob: the reference's Gibbs driver as a CALLER of the batched MCMC hot path.

    BasicGibbsJob(model, dpjob, mcrange, v0; outopts)              src/jobs/BasicGibbsJob.jl:3-148
    run(job)                                                        src/jobs/BasicGibbsJob.jl:201-231
      for i = 1:nsteps
        iterate!(job)                                               :185-199
          for every dependent variable (parameters and transformations, in vertex order)
            parameter with a BasicMCJob:  run(dpjob); dpstate = dpjob.pstate
            transformation:               transform!(dpstate)
        if i in postrange: count += 1; save(job, count)             :170-183 (copy! of every dependent's state)
        reset(job)                                                  :158-168 -> reset(dpjob) for every dpjob

Each `run(dpjob)` is one klb_job_run (all the inner job's transitions for all chains in one launch) and each
`reset(dpjob)` one klb_job_reset: the sweep loop drives the existing C ABI and nothing else.  The RNG counter of a
dpjob keeps advancing across sweeps (the reference's global RNG does too), the tuner record and the inner output
cursor restart every sweep, the chain state persists.

Scope.  A parameter needs a dpjob: sampling a parameter from a `setpdf` closure (Distributions.jl, host side) is
outside the device path.  Device targets take their hyper-parameters when the dpjob is built; re-binding them between
sweeps from the other blocks' fresh values is not done here (the reference would then keep a stale cached log-target
and gradient in dpjob.pstate for the first transition of the next sweep -- behaviour that cannot be pinned without
running it), so the blocks are conditionally independent given the fixed hyper-parameters; transformations see the
fresh states of every block."""
import numpy as np

from .api import BasicContMuvParameter, BasicContMuvParameterNState, BasicMCJob, BasicMCRange

__all__ = ["Transformation", "BasicGibbsJob"]


class Transformation:
    """Transformation(key; transform): a deterministic vertex, recomputed in every sweep from the current values of the
    model's vertices (src/variables/variables.jl:97-123).  `transform(values)` receives a dict key -> array
    ((nchains, dim) for parameters) and returns an (nchains, k) array."""

    def __init__(self, key, transform, index=0):
        self.key, self.transform, self.index = key, transform, index


class BasicGibbsJob:
    def __init__(self, model, dpjob, mcrange, v0, outopts=None, verbose=False):
        if not isinstance(mcrange, BasicMCRange):
            raise TypeError("mcrange must be a BasicMCRange")
        self.model, self.range, self.verbose = model, mcrange, verbose
        # dpindex: parameters and transformations, in model-vertex order            BasicGibbsJob.jl:95
        self.dependent = [v for v in model.vertices if isinstance(v, (BasicContMuvParameter, Transformation))]
        if not self.dependent:
            raise ValueError("The model has neither parameters nor transformations, but at least one of them is required "
                             "in a BasicGibbsJob")
        self.dpjob = {}
        for v in self.dependent:
            j = dpjob.get(v.key)
            if isinstance(v, BasicContMuvParameter):
                if not isinstance(j, BasicMCJob):
                    raise TypeError("parameter %r needs a BasicMCJob: sampling from a setpdf closure is host-side and outside "
                                    "the device path" % (v.key,))
                self.dpjob[v.key] = j
        chains = {j.nchains for j in self.dpjob.values()}
        if len(chains) > 1:
            raise AssertionError("all dpjobs must carry the same number of chains, got %s" % sorted(chains))
        self.nchains = chains.pop() if chains else 1
        self.vstate = {k: np.asarray(x, dtype=np.float64) if not isinstance(x, (int, float)) else x for k, x in v0.items()}
        oo = outopts or {}
        self.outopts = {v.key: dict({"destination": "nstate", "monitor": ["value"]}, **oo.get(v.key, {})) for v in self.dependent}
        self.output = {}
        self.count = 0

    def _states(self):
        return dict(self.vstate)

    def iterate(self, fetch):
        """iterate!(job)        src/jobs/BasicGibbsJob.jl:185-199"""
        for v in self.dependent:
            if isinstance(v, BasicContMuvParameter):
                j = self.dpjob[v.key]
                j.run()                                                   # run(job.dpjob[i])
                if fetch:
                    self.vstate[v.key] = np.atleast_2d(j.pstate_value)    # job.dpstate[i] = job.dpjob[i].pstate
            elif fetch:
                self.vstate[v.key] = np.asarray(v.transform(self._states()), dtype=np.float64)   # transform!(dpstate)

    def reset(self):
        """reset(job): reset(dpjob) for every dpjob        src/jobs/BasicGibbsJob.jl:158-168 (resetpstate = false)"""
        for j in self.dpjob.values():
            j.reset()

    def run(self):
        """run(job)        src/jobs/BasicGibbsJob.jl:201-231"""
        r = self.range
        npost = r.npoststeps
        has_tr = any(isinstance(v, Transformation) for v in self.dependent)
        for i in range(1, r.nsteps + 1):
            save = i > r.burnin and (i - r.burnin - 1) % r.thinning == 0      # in(i, postrange)
            self.iterate(fetch=save or has_tr)
            if save:
                self.count += 1
                for v in self.dependent:                                       # save(job, count): copy! per dependent
                    if self.outopts[v.key]["destination"] != "nstate":
                        continue
                    val = np.atleast_2d(self.vstate[v.key])
                    ns = self.output.get(v.key)
                    if ns is None:
                        ns = self.output[v.key] = BasicContMuvParameterNState(val.shape[-1], npost)
                        ns.value = np.empty((val.shape[0], npost, val.shape[-1]))
                    ns.value[:, self.count - 1, :] = val
            self.reset()
        return self

    def output_dict(self):
        """Dict(job): key -> NState of every dependent variable        src/jobs/BasicGibbsJob.jl:292-300"""
        return dict(self.output)
