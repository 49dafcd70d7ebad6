#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Klara.jl MCMC hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json configs[2], "C3"): HMC(leapstep 0.05, nleaps 10) on the 1024-dim isotropic
Gaussian  logtarget(z) = -z.z, 65 536 independent chains, BasicMCRange(nsteps=200, burnin=100),
monitor [:value, :logtarget], diagnostics [:accept]  (SURVEY.md section 8d).  One "step" of this
benchmark = one complete run(job) of that BasicMCJob over all chains (reset + 200 transitions =
2000 leapfrog steps per chain, 100 stored samples per chain).  Metric: leapfrog steps per second,
whole job, all GPUs.  Chains are sharded over ranks (strong scaling: the 65 536 chains are fixed);
the RNG is keyed by the global chain index so results do not depend on the number of GPUs; one
NCCL all-gather of the final states closes every step.

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240925
NCHAINS, DIM, NLEAPS, LEAPSTEP, NSTEPS, BURNIN = 65536, 1024, 10, 0.05, 200, 100
METRIC, UNIT = "leapfrog_steps_per_sec", "leapfrog-steps/s"
WORKLOAD = ("C3: HMC(leapstep=0.05, nleaps=10), isotropic Gaussian logtarget -z.z, 65536 chains x 1024 dim, "
            "BasicMCRange(nsteps=200, burnin=100), monitor value+logtarget, diagnostics accept, fp64, "
            "arith=reference (un-fused)")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle (line-by-line restatement of the reference loop) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_leg(nthreads=None, target_seconds=12.0):
    """Times the reference algorithm (oracle/klb_oracle.c; Klara.jl itself needs Julia 0.6, absent) on a
    bounded sample of the same workload: same d, L, nsteps/burnin/monitor, fewer chains."""
    from oracle import oracle as O
    O.build()
    nthreads = nthreads or os.cpu_count() or 1
    cfgp = dict(step=LEAPSTEP, nleaps=NLEAPS, monitor=3, diagnostics=1, seed=SEED, nthreads=nthreads)
    def timed(nchains):
        x0 = np.stack([O.normals(SEED, c, 0, DIM) for c in range(nchains)])
        cfg = O.make_config(O.HMC, O.ISO, nchains, DIM, NSTEPS, BURNIN, **cfgp)
        t = time.perf_counter()
        res = O.run(cfg, x0)
        return time.perf_counter() - t, res
    # probe with two chains per thread, then size the sample so the timed run lasts ~target_seconds
    nchains = 2 * nthreads
    dt, res = timed(nchains)
    if dt < 0.6 * target_seconds:
        nchains = int(min(65536, max(nchains, nchains * target_seconds / max(dt, 1e-3))) // nthreads * nthreads)
        dt, res = timed(nchains)
    lf = nchains * NLEAPS * NSTEPS
    return {"value": lf / dt, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "%d of 65536 chains, full nsteps=%d (burnin %d), d=%d, L=%d, OpenMP over chains; %.1f s"
                      % (nchains, NSTEPS, BURNIN, DIM, NLEAPS, dt),
            "seconds": dt, "accept_rate": float(res["accept"].mean())}, res


def run_reference(args, rank):
    if rank != 0:
        return
    t_all = time.perf_counter()
    budget = float(os.environ.get("KLB_BENCH_CPU_SECONDS", "20"))     # total CPU time of the timed steps
    for _ in range(args.warmup):
        cpu_leg(target_seconds=min(1.0, budget))
    vals = []
    for _ in range(args.steps):
        cb, _ = cpu_leg(target_seconds=max(min(2.0, budget), budget / max(1, args.steps)))
        vals.append(cb)
    best = max(vals, key=lambda c: c["value"])
    mean_v = float(np.mean([c["value"] for c in vals]))
    total = time.perf_counter() - t_all
    out = {"impl": "reference", "metric": METRIC, "value": mean_v, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([c["seconds"] for c in vals])),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "note": "reference algorithm (CPU oracle = C restatement of Klara.jl's "
                      "BasicMCJob loop; Klara.jl itself needs Julia 0.6, not installable offline) on a bounded "
                      "chain subset, all host threads; an upper bound on Klara.jl's own allocating, dynamically "
                      "dispatched loop"},
           "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": mean_v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": total}
    out["cpu_baseline"]["value"] = mean_v
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import klara_b200 as K
    L = K._lib
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    assert NCHAINS % world == 0
    nloc = NCHAINS // world
    off = rank * nloc
    arith = args.arith

    # synthetic initial state, generated on the device from the Philox streams (seed, chain, t=0) --
    # x0[c] = N(0, I); the resident-HBM leg starts from device memory, the e2e leg from pinned host memory
    x0_host = np.empty((nloc, DIM))
    lib = L.lib()
    hx = C.c_void_p()
    L.check(lib.klb_host_alloc(C.byref(hx), x0_host.nbytes))
    x0_pin = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_double)), shape=(nloc, DIM))
    rng = np.random.default_rng(SEED + rank)
    x0_pin[:] = rng.standard_normal((nloc, DIM))

    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    model = K.likelihood_model(p, False)
    job = K.BasicMCJob(model, K.HMC(LEAPSTEP, NLEAPS), K.BasicMCRange(nsteps=NSTEPS, burnin=BURNIN), {"p": x0_pin},
                       outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]},
                       seed=SEED, arith=arith, device=local_rank, chain_offset=off)
    plan = job.plan()
    stream = torch.cuda.ExternalStream(lib.klb_job_stream(job._h), device=torch.device("cuda", local_rank))

    # zero-copy torch view of the library-owned final state for the closing NCCL all-gather
    sp, snb = job.device_ptr(L.OUT_STATE)

    class _Iface:
        __cuda_array_interface__ = {"shape": (nloc, DIM), "typestr": "<f8", "data": (sp, False), "version": 2}
    state_t = torch.as_tensor(_Iface(), device=torch.device("cuda", local_rank))
    gathered = torch.empty((world, nloc, DIM), dtype=torch.float64, device=state_t.device) if world > 1 else None
    # The closing all-gather of run k reads a snapshot of the final states (a 64 MiB device copy at N=8) on its own
    # stream, so that it overlaps the kernel of run k+1 instead of delaying it; everything is joined before the
    # closing event of the timed region.
    snapshot = torch.empty_like(state_t) if world > 1 else None
    comm_stream = torch.cuda.Stream(device=state_t.device) if world > 1 else None
    gather_done = torch.cuda.Event() if world > 1 else None

    def one_step():
        job.reset()
        job.run_async()
        if world > 1:
            with torch.cuda.stream(stream):
                stream.wait_event(gather_done)          # the previous gather has read the snapshot
                snapshot.copy_(state_t, non_blocking=True)
            comm_stream.wait_stream(stream)
            with torch.cuda.stream(comm_stream):
                dist.all_gather_into_tensor(gathered.view(-1), snapshot.view(-1))
                gather_done.record(comm_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        gather_done.record(comm_stream)
    for _ in range(args.warmup):
        one_step()
    barrier()
    launches_w = job.launches
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
        if args.per_step_sync:
            job.sync()
    if world > 1:
        stream.wait_stream(comm_stream)                 # the last all-gather belongs to the timed region
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    last_kernel_ms = job.last_run_ms
    clocks = sampler.finish() if sampler else None
    launches_timed = job.launches - launches_w

    # -------- end-to-end leg: host buffers in, host buffers out, through the public API
    lt_host = np.empty((nloc, NSTEPS - BURNIN))
    acc_host = np.empty((nloc, NSTEPS - BURNIN), dtype=np.uint8)
    hs = C.c_void_p()
    L.check(lib.klb_host_alloc(C.byref(hs), x0_host.nbytes))
    st_pin = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_double)), shape=(nloc, DIM))
    hl = C.c_void_p()
    L.check(lib.klb_host_alloc(C.byref(hl), lt_host.nbytes + acc_host.nbytes))
    h2d = x0_host.nbytes
    d2h = x0_host.nbytes + lt_host.nbytes + acc_host.nbytes

    e2e_fields = (L.KlbHostField * 3)()
    for q, (fld, ptr, nb) in enumerate(((L.OUT_STATE, hs.value, x0_host.nbytes), (L.OUT_LOGTARGET, hl.value, lt_host.nbytes),
                                        (L.OUT_ACCEPT, hl.value + lt_host.nbytes, acc_host.nbytes))):
        e2e_fields[q].field, e2e_fields[q].host_dst, e2e_fields[q].nbytes = fld, ptr, nb

    def e2e_step():
        if args.e2e_serial:                                        # the three blocking calls, one after the other
            L.check(lib.klb_job_set_state(job._h, hx))             # H2D x0 + initialize! + tuner reset
            L.check(lib.klb_job_run(job._h))
            L.check(lib.klb_job_output(job._h, L.OUT_STATE, hs, x0_host.nbytes))
            L.check(lib.klb_job_output(job._h, L.OUT_LOGTARGET, hl, lt_host.nbytes))
            L.check(lib.klb_job_output(job._h, L.OUT_ACCEPT, C.c_void_p(hl.value + lt_host.nbytes), acc_host.nbytes))
        else:                                                      # the same work as one pipelined call
            L.check(lib.klb_job_run_host(job._h, hx, e2e_fields, 3, args.e2e_slices))

    e2e_step()
    barrier()
    te = time.perf_counter()
    nrep = max(1, min(args.steps, 3))
    for _ in range(nrep):
        e2e_step()
    barrier()
    e2e_wall = (time.perf_counter() - te) / nrep
    acc_rate = float(np.ctypeslib.as_array(C.cast(C.c_void_p(hl.value + lt_host.nbytes), C.POINTER(C.c_uint8)),
                                           shape=(nloc, NSTEPS - BURNIN)).mean())

    # -------- optional: end-to-end with EVERY monitored field copied to the host (50 GiB of values at N=1)
    e2e_full = None
    if args.e2e_full:
        vbytes = nloc * (NSTEPS - BURNIN) * DIM * 8
        hv = C.c_void_p()
        L.check(lib.klb_host_alloc(C.byref(hv), vbytes))

        def full_step():
            e2e_step()
            L.check(lib.klb_job_output(job._h, L.OUT_VALUE, hv, vbytes))
        full_step()
        barrier()
        tf = time.perf_counter()
        full_step()
        barrier()
        e2e_full = {"ms_per_step": (time.perf_counter() - tf) * 1e3, "d2h_bytes_per_step": (d2h + vbytes) * world}
        lib.klb_host_free(hv)

    # -------- effective sample size of the stored chains, on the device
    tq = time.perf_counter()
    L.check(lib.klb_job_ess(job._h, None))
    ess_ms = (time.perf_counter() - tq) * 1e3
    ep, enb = job.device_ptr(L.OUT_ESS)

    class _IfaceE:
        __cuda_array_interface__ = {"shape": (nloc, DIM), "typestr": "<f8", "data": (ep, False), "version": 2}
    ess_t = torch.as_tensor(_IfaceE(), device=torch.device("cuda", local_rank))
    ess_stats = torch.stack([ess_t.sum(), ess_t.min(), torch.isfinite(ess_t).all().double()])
    ess_stats[1] = -ess_stats[1]
    if world > 1:
        s_ = ess_stats[:1].clone(); dist.all_reduce(s_); ess_stats[0] = s_[0]
        m_ = ess_stats[1:2].clone(); dist.all_reduce(m_, op=dist.ReduceOp.MAX); ess_stats[1] = m_[0]
    ess_sum, ess_min = float(ess_stats[0]), -float(ess_stats[1])

    # -------- reduce over ranks: max time
    t_dev = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3, last_kernel_ms], dtype=torch.float64,
                         device=state_t.device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms, kernel_ms = [float(v) for v in t_dev.cpu()]

    lf_per_step = NCHAINS * NLEAPS * NSTEPS            # leapfrog steps in one bench step, all ranks
    value = lf_per_step * args.steps / (dev_ms * 1e-3)
    out = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        # algorithmic bytes per launch: per transition read x + write x (16 d) for every chain,
        # plus (8 d + 9) per stored sample (value + logtarget + accept flag); this rank's launch covers nloc chains
        bytes_launch = nloc * (NSTEPS * 16 * DIM + (NSTEPS - BURNIN) * (8 * DIM + 9))
        achieved = bytes_launch / (kernel_ms * 1e-3) / 1e9
        # fp64 issue roofline: 5 d un-fused ops per leapfrog step (+ per-transition overhead ignored)
        fp64_ops = nloc * NSTEPS * (NLEAPS * (5 if arith == "reference" else 3) * DIM)
        cb = None
        if world == 1 and not args.no_cpu:
            cb, _ = cpu_leg(target_seconds=12.0)
            cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_gpu": nloc, "arith": arith,
                       "l2": "inputs larger than L2 (state 512 MiB + 50 GiB of samples per step at N=1)",
                       "nv": plan.nv, "regs_per_thread": plan.regs_per_thread, "blocks_per_sm": plan.blocks_per_sm,
                       "accept_rate": acc_rate, "timing": "CUDA events on the job stream, max over ranks",
                       "wall_ms_per_step": wall_ms / args.steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this launch at N=1 (ncu --set full,
                         # profiles/r1_summary.md capture G), scaled to this rank's chains
                         "traffic": 54.89e9 * nloc / NCHAINS if arith == "reference" else None,
                         "traffic_source": "profiles/r1_kernel_metrics.csv column M_ws_shipping_bench_launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                         "peak_source": peak_src,
                         "kernel": "klb_hmc_ws_kernel<TgtIso, NV=16> (warp-specialised: 4 consumer + 4 producer warps per CTA)", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": bytes_launch,
                         "fp64": {"achieved_tops": fp64_ops / (kernel_ms * 1e-3) / 1e12,
                                  "peak_tops": 148 * 64 * 1.965e-3,
                                  "note": "leapfrog fp64 instructions only (5 d per step, un-fused: DADD/DMUL count 1 each) "
                                          "against 148 SMs x 64 lanes x 1.965 GHz; the kernel is fp64-issue bound, not HBM bound "
                                          "(ncu, column M: fp64 pipe 61.9 % busy, issue slots 57.7 %, DRAM 10.0 %): profiles/r1_summary.md"}},
            "cpu_baseline": cb,
            "e2e": {"value": lf_per_step / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms,
                    "call": "klb_job_set_state + klb_job_run + klb_job_output x3 (serial)" if args.e2e_serial
                            else "klb_job_run_host (chain slices pipelined over streams)",
                    "note": "host x0 in (pinned), final state + logtarget chain + accept flags out (pinned), through the "
                            "C ABI; the 50 GiB of monitored values stay in HBM (output(job) copies them on request)"},
            "ess": {"mean_ess_per_coordinate": ess_sum / (NCHAINS * DIM), "min_ess": ess_min, "samples_per_chain": NSTEPS - BURNIN,
                    "independent_samples_per_sec": (ess_sum / DIM) / (dev_ms / args.steps * 1e-3),
                    "ess_kernel_ms": ess_ms,
                    "note": "ess(chain, :imse) per coordinate on the device (klb_job_ess); independent samples/s = "
                            "sum over chains of the coordinate-mean ESS / device time of one run"},
            "e2e_full_output": None if e2e_full is None else dict(
                e2e_full, value=lf_per_step / (e2e_full["ms_per_step"] * 1e-3), unit=UNIT,
                note="as e2e, plus klb_job_output(KLB_OUT_VALUE): all monitored samples copied to pinned host memory"),
            "gpu_launches": int(launches_timed),
            "clocks": clocks,
        }
    job.close()
    lib.klb_host_free(hx); lib.klb_host_free(hs); lib.klb_host_free(hl)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arith", default="reference", choices=["reference", "fma"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--per-step-sync", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e leg through the three blocking calls (set_state, run, output) instead of klb_job_run_host")
    ap.add_argument("--e2e-slices", type=int, default=0, help="chain slices of the pipelined e2e call (0 = library default)")
    ap.add_argument("--e2e-full", action="store_true", help="also time an end-to-end step that copies every monitored sample to the host")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus == 1 and world == 1:
            pass
        else:
            raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
