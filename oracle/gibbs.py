"""gibbs.py -- CPU ORACLE of the Gibbs sweep driver.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates run(job::BasicGibbsJob) (src/jobs/BasicGibbsJob.jl:201-231) with iterate! (:185-199), save (:170-183) and
reset (:158-168) for dependent variables that are parameters sampled by a BasicMCJob, or transformations:

    for i = 1:nsteps
      for every dependent variable j (vertex order):
        parameter:       run(dpjob[j])   -> oracle.run on the block's config, continuing from its persistent state:
                                            value, cached logtarget (initialized = true), RNG counter t0 advanced by the
                                            inner job's nsteps per sweep
                         dpstate[j] = dpjob[j].pstate
        transformation:  transform!(dpstate[j])
      if i in postrange: count += 1; copy!(output[j], dpstate[j], count) for every j
      reset(dpjob[j]) for every dpjob:  tuner record <- tuner_state, inner output cursor <- 0; pstate persists
"""
import numpy as np

from . import oracle as O


def run_gibbs(blocks, transforms, order, nsteps, burnin=0, thinning=1):
    """blocks: key -> dict(cfg=OrcConfig of the inner BasicMCJob, x0=(nchains, dim), tparams, sigma)
    transforms: key -> callable(values dict) -> (nchains, k) array;  order: the dependent variables' keys in vertex order.
    Returns key -> (nchains, npost, dim) array of saved values."""
    state = {k: dict(x=np.array(b["x0"], dtype=np.float64), lt=None, t0=0) for k, b in blocks.items()}
    values = {k: s["x"] for k, s in state.items()}
    npost = O.npoststeps(burnin, thinning, nsteps)
    out, count = {}, 0
    for i in range(1, nsteps + 1):
        for key in order:                                        # iterate!(job)
            if key in blocks:
                b, s = blocks[key], state[key]
                cfg = b["cfg"]
                cfg.t0 = s["t0"]
                # reset(dpjob) of the previous sweep left a fresh tuner record: oracle.run builds it when tune is None
                r = O.run(cfg, s["x"], b.get("tparams"), b.get("sigma"), logtarget=s["lt"])
                s["x"], s["lt"], s["t0"] = r["x"], r["logtarget_state"], s["t0"] + cfg.nsteps
                values[key] = s["x"]
            else:
                values[key] = np.asarray(transforms[key](dict(values)), dtype=np.float64)
        if i > burnin and (i - burnin - 1) % thinning == 0:     # in(i, postrange) -> save(job, count)
            count += 1
            for key in order:
                v = np.atleast_2d(values[key])
                if key not in out:
                    out[key] = np.empty((v.shape[0], npost, v.shape[-1]))
                out[key][:, count - 1, :] = v
    return out
