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
the initial value and the RNG are keyed by the GLOBAL chain index (Philox streams (seed, chain, t)),
so inputs and results do not depend on the number of GPUs; one all-gather of the final states and
tuner records closes every step (copy engines over NVLink by default, --gather nccl for NCCL).

Besides the timed legs the run VERIFIES itself (`parity`): a fresh run from the synthetic initial value
is compared, on a chain subset, with the CPU oracle (accept/reject sequence, log-targets, values, final
state: bit for bit), at N > 1 chains of the OTHER ranks' shards are recomputed on rank 0 and compared with
the all-gathered result (G-invariance), and a checksum of all final states is printed that must be the same
for every N.  `configs` carries the other BASELINE.json configurations (C2, C4, C5, MH) at their stated sizes and HMC at dim 4096.

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
NPOST = NSTEPS - BURNIN
METRIC, UNIT = "leapfrog_steps_per_sec", "leapfrog-steps/s"
WORKLOAD = ("C3: HMC(leapstep=0.05, nleaps=10), isotropic Gaussian logtarget -z.z, 65536 chains x 1024 dim, "
            "BasicMCRange(nsteps=200, burnin=100), monitor value+logtarget, diagnostics accept, fp64, "
            "arith=reference (un-fused)")
VERIFY_T = 0            # the verification run replays transitions 1 .. NSTEPS of the streams (what a fresh job does)


def job_config(arith):
    """`config` of the JSON line: what the workload IS.  The same dict in both arms (the reference arm runs on this arm's
    config); what a particular run looked like goes under `details`."""
    return {"workload": WORKLOAD, "arith": arith,
            "l2": "inputs larger than L2 (state 512 MiB + 50 GiB of samples per step at N=1)",
            "x0": "Philox stream (seed, global chain, transition 0)",
            "rng": "Philox4x32-7 counter streams + 256-layer ziggurat (DESIGN.md section 3)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic():
    """dram bytes per C3 launch from the committed ncu capture of this round (profiles/r2_traffic.json)"""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None, "no capture committed"
    with open(p) as fh:
        t = json.load(fh)
    return t, "profiles/r2_traffic.json (%s)" % t.get("source", "ncu --set full")


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
    bounded sample of the same workload: same d, L, nsteps/burnin/monitor, the FIRST chains of the job from
    their synthetic initial value.  Its outputs are what `parity` compares the CUDA path with."""
    from oracle import oracle as O
    O.build()
    nthreads = nthreads or os.cpu_count() or 1
    cfgp = dict(step=LEAPSTEP, nleaps=NLEAPS, diagnostics=1, seed=SEED, nthreads=nthreads)

    def timed(nchains, monitor):
        x0 = np.stack([O.normals(SEED, c, 0, DIM) for c in range(nchains)])
        cfg = O.make_config(O.HMC, O.ISO, nchains, DIM, NSTEPS, BURNIN, monitor=monitor, **cfgp)
        t = time.perf_counter()
        res = O.run(cfg, x0)
        return time.perf_counter() - t, res
    # probe with two chains per thread (values monitored: these chains carry the value comparison of `parity`), then
    # size the sample so the timed run lasts ~target_seconds.  The large sample monitors the log-target and the accept
    # flags only: its 100 x 1024 values per chain would need 0.8 MB of host memory each (the copies it skips are
    # ~1 % of a chain's work, in the CPU's favour).
    nprobe = 2 * nthreads
    dt, res = timed(nprobe, 3)
    nchains = nprobe
    if dt < 0.6 * target_seconds:
        nchains = int(min(32768, max(nprobe, nprobe * target_seconds / max(dt, 1e-3))) // nthreads * nthreads)
        dt, big = timed(nchains, 2)
        big["value"] = res["value"]                      # first `nprobe` chains: the same chains, the same streams
        res = big
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
           "config": job_config(args.arith),
           "details": {"note": "reference algorithm (CPU oracle = C restatement of Klara.jl's "
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
def _view(torch, ptr, shape, dtype, device):
    class _Iface:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": {"f8": "<f8", "u1": "|u1", "i8": "<i8"}[dtype],
                                    "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Iface(), device=device)


class Pinned:
    """pinned host array (klb_host_alloc)"""

    def __init__(self, L, shape, dtype=np.float64):
        self.L, self.h = L, C.c_void_p()
        nb = int(np.prod(shape)) * np.dtype(dtype).itemsize
        L.check(L.lib().klb_host_alloc(C.byref(self.h), nb))
        ct = {np.dtype(np.float64): C.c_double, np.dtype(np.uint8): C.c_uint8}[np.dtype(dtype)]
        self.a = np.ctypeslib.as_array(C.cast(self.h, C.POINTER(ct)), shape=shape)
        self.nbytes = nb

    def free(self):
        if self.h:
            self.L.lib().klb_host_free(self.h)
            self.h, self.a = None, None


def run_gpu(args, rank, world, local_rank):
    import torch
    import klara_b200 as K
    L = K._lib
    lib = L.lib()
    dist = None
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lo, hi = K.distributed.shard_range(NCHAINS, rank, world)
    nloc = hi - lo
    arith = args.arith

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    # -------- the job: synthetic initial value generated on the device from the Philox streams (seed, chain, 0)
    p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
    model = K.likelihood_model(p, False)
    job = K.BasicMCJob(model, K.HMC(LEAPSTEP, NLEAPS), K.BasicMCRange(nsteps=NSTEPS, burnin=BURNIN),
                       {"p": K.SyntheticNormal(nloc, DIM)},
                       outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept"]},
                       seed=SEED, arith=arith, device=local_rank, chain_offset=lo)
    plan = job.plan()
    stream = torch.cuda.ExternalStream(lib.klb_job_stream(job._h), device=dev)
    x0_pin = Pinned(L, (nloc, DIM))                     # the same initial value in pinned host memory (e2e leg)
    L.check(lib.klb_job_output(job._h, L.OUT_STATE, x0_pin.h, x0_pin.nbytes))
    sp, _ = job.device_ptr(L.OUT_STATE)
    state_t = _view(torch, sp, (nloc, DIM), "f8", dev)

    # -------- the closing all-gather
    gather, gathered_t, nccl, gather_note = None, None, None, None
    if world > 1 and args.gather == "p2p":
        g = C.c_void_p()
        L.check(lib.klb_gather_create(job._h, world, rank, NCHAINS, 0, C.byref(g)))
        h = C.create_string_buffer(L.GATHER_HANDLE_BYTES)
        L.check(lib.klb_gather_handle(g, h))
        mine = torch.frombuffer(bytearray(h.raw), dtype=torch.uint8).to(dev)
        allh = torch.empty(world * L.GATHER_HANDLE_BYTES, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine)         # plumbing: the IPC handles travel over NCCL
        raw = bytes(allh.cpu().numpy())
        ok = torch.ones(1, device=dev)
        try:
            L.check(lib.klb_gather_connect(g, C.create_string_buffer(raw, len(raw))))
        except L.KlaraError as e:                       # e.g. CUDA IPC not permitted between these processes
            ok.zero_()
            gather_note = str(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)       # every rank takes the same path
        if float(ok) > 0:
            gp, gnb = C.c_void_p(), C.c_int64()
            L.check(lib.klb_gather_device_ptr(g, L.OUT_STATE, C.byref(gp), C.byref(gnb)))
            gathered_t = _view(torch, gp.value, (NCHAINS, DIM), "f8", dev)
            gather = g
        else:
            lib.klb_gather_destroy(g)
    if world > 1 and gather is None:
        # NCCL all-gather of a snapshot of the final states on its own stream (round 1's path: SM-based kernel)
        nccl = {"out": torch.empty((NCHAINS, DIM), dtype=torch.float64, device=dev), "snap": torch.empty_like(state_t),
                "stream": torch.cuda.Stream(device=dev), "done": torch.cuda.Event()}
        nccl["done"].record(nccl["stream"])
        gathered_t = nccl["out"]

    def push():
        if gather is not None:
            L.check(lib.klb_gather_push_async(gather))
        elif nccl is not None:
            with torch.cuda.stream(stream):
                stream.wait_event(nccl["done"])         # the previous gather has read the snapshot
                nccl["snap"].copy_(state_t, non_blocking=True)
            nccl["stream"].wait_stream(stream)
            with torch.cuda.stream(nccl["stream"]):
                dist.all_gather_into_tensor(nccl["out"].view(-1), nccl["snap"].view(-1))
                nccl["done"].record(nccl["stream"])

    def join():                                         # the job stream waits for the last all-gather
        if gather is not None:
            L.check(lib.klb_gather_join(gather))
        elif nccl is not None:
            stream.wait_stream(nccl["stream"])

    def one_step():
        job.reset()
        job.run_async()
        push()

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
    join()                                              # the last all-gather belongs to the timed region
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    last_kernel_ms = job.last_run_ms
    clocks = sampler.finish() if sampler else None
    launches_timed = job.launches - launches_w

    # -------- end-to-end leg: host buffers in, host buffers out, through the public API
    st_pin, lt_pin, ac_pin = Pinned(L, (nloc, DIM)), Pinned(L, (nloc, NPOST)), Pinned(L, (nloc, NPOST), np.uint8)
    h2d = x0_pin.nbytes
    d2h = st_pin.nbytes + lt_pin.nbytes + ac_pin.nbytes

    def fields(*extra):
        spec = [(L.OUT_STATE, st_pin), (L.OUT_LOGTARGET, lt_pin), (L.OUT_ACCEPT, ac_pin)] + list(extra)
        arr = (L.KlbHostField * len(spec))()
        for q, (fld, buf) in enumerate(spec):
            arr[q].field, arr[q].host_dst, arr[q].nbytes = fld, buf.h.value, buf.nbytes
        return arr, len(spec)
    e2e_fields, ne2e = fields()

    def e2e_step():
        if args.e2e_serial:                                        # the three blocking calls, one after the other
            L.check(lib.klb_job_set_state(job._h, x0_pin.h))      # H2D x0 + initialize! + tuner reset
            L.check(lib.klb_job_run(job._h))
            L.check(lib.klb_job_output(job._h, L.OUT_STATE, st_pin.h, st_pin.nbytes))
            L.check(lib.klb_job_output(job._h, L.OUT_LOGTARGET, lt_pin.h, lt_pin.nbytes))
            L.check(lib.klb_job_output(job._h, L.OUT_ACCEPT, ac_pin.h, ac_pin.nbytes))
        else:                                                      # the same work as one pipelined call
            L.check(lib.klb_job_run_host(job._h, x0_pin.h, e2e_fields, ne2e, args.e2e_slices))

    e2e_step()
    barrier()
    te = time.perf_counter()
    nrep = max(1, min(args.steps, 3))
    for _ in range(nrep):
        e2e_step()
    barrier()
    e2e_wall = (time.perf_counter() - te) / nrep
    acc_rate = float(ac_pin.a.mean())

    # -------- end-to-end with EVERY monitored field copied to the host: what output(job) returns in the reference
    # (50 GiB of values at N = 1).  The values ride in the same pipelined call, slice by slice.
    e2e_full = None
    if not args.no_e2e_full:
        try:
            v_pin = Pinned(L, (nloc, NPOST, DIM))
            full_fields, nfull = fields((L.OUT_VALUE, v_pin))
            L.check(lib.klb_job_run_host(job._h, x0_pin.h, full_fields, nfull, args.e2e_slices))
            barrier()
            tf = time.perf_counter()
            L.check(lib.klb_job_run_host(job._h, x0_pin.h, full_fields, nfull, args.e2e_slices))
            barrier()
            e2e_full = {"ms_per_step": (time.perf_counter() - tf) * 1e3, "d2h_bytes_per_step": (d2h + v_pin.nbytes) * world,
                        "h2d_bytes_per_step": h2d * world}
            v_pin.free()
        except L.KlaraError as e:                       # not enough pinned host memory on this box
            e2e_full = {"unavailable": str(e)}

    # -------- effective sample size of the stored chains, on the device
    job.reset(); job.run()
    L.check(lib.klb_job_ess(job._h, None))             # first call allocates the result buffer
    torch.cuda.synchronize()
    tq = time.perf_counter()
    L.check(lib.klb_job_ess(job._h, None))
    ess_ms = (time.perf_counter() - tq) * 1e3
    ep, _ = job.device_ptr(L.OUT_ESS)
    ess_t = _view(torch, ep, (nloc, DIM), "f8", dev)
    ess_stats = torch.stack([ess_t.sum(), -ess_t.min()])
    if world > 1:
        s_ = ess_stats[:1].clone(); dist.all_reduce(s_); ess_stats[0] = s_[0]
        m_ = ess_stats[1:2].clone(); dist.all_reduce(m_, op=dist.ReduceOp.MAX); ess_stats[1] = m_[0]
    ess_sum, ess_min = float(ess_stats[0]), -float(ess_stats[1])

    # -------- verification run: a fresh job's first run (synthetic x0, transitions 1..nsteps), all ranks, gathered
    job.reset_synthetic()
    job.seek(VERIFY_T)
    job.run_async()
    push()
    join()
    job.sync()
    if gather is not None:
        L.check(lib.klb_gather_sync(gather))
    barrier()
    full_state = gathered_t if world > 1 else state_t          # (NCHAINS, DIM) on every rank
    words = full_state.view(torch.int64).reshape(-1)
    w = (torch.arange(words.numel(), device=dev, dtype=torch.int64) % 65521) + 1
    checksum = "%016x-%016x" % (int(words.sum().item()) & (2 ** 64 - 1), int((words * w).sum().item()) & (2 ** 64 - 1))
    parity = None
    if rank == 0:
        parity = verify(args, torch, K, L, job, nloc, lo, world, full_state, dev)
        parity["final_state_checksum"] = checksum
        parity["note"] = ("fresh run from the synthetic initial value: chains 0..n-1 against the CPU oracle (bit for bit; "
                          "max_rel_* are 0 when the bits agree); at N > 1 chains of the other ranks' shards recomputed on "
                          "rank 0 against the all-gathered states; final_state_checksum covers all 65 536 final states and "
                          "is the same for every N")
    cb = parity.pop("_cpu_baseline", None) if parity else None

    # -------- reduce over ranks: max time
    dev_ms, wall_ms, e2e_ms, kernel_ms = allmax([dev_ms, wall * 1e3, e2e_wall * 1e3, last_kernel_ms])
    if e2e_full and "ms_per_step" in e2e_full:
        e2e_full["ms_per_step"] = allmax([e2e_full["ms_per_step"]])[0]

    # -------- free the C3 job, then the other BASELINE configurations at their stated sizes
    if gather is not None:
        barrier()
        L.check(lib.klb_gather_disconnect(gather))      # every rank unmaps its peers' buffers ...
        barrier()
        lib.klb_gather_destroy(gather)                  # ... before any rank frees its own
    job.close()
    for b in (x0_pin, st_pin, lt_pin, ac_pin):
        b.free()
    del state_t, ess_t, full_state, words, w, gathered_t
    # fp64 issue-rate roofline: one warp instruction per 2 cycles per scheduler = 64 results per cycle and SM
    # (tools/fp64_pipe_test.cu measured 2.005 cycles per instruction, profiles/r1_summary.md), at the part's maximum SM
    # clock.  A stream of independent DFMA timed with CUDA events (klb_device_peak) lands ~8 % below it (all-DFMA load),
    # so the nominal-at-max-clock figure is the stricter denominator; both are printed under `peaks`.
    fp64_stream = K.device_peak("fp64", local_rank)
    sm_mhz_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp64_peak = max(fp64_stream, sms * 64 * sm_mhz_max * 1e6)
    dmma_peak = K.device_peak("dmma", local_rank)
    hbm_peak, peak_src = measured_peaks()
    configs = None
    if not args.no_configs:
        configs = other_configs(args, torch, K, L, rank, world, local_rank, allmax, fp64_peak, dmma_peak, hbm_peak)

    lf_per_step = NCHAINS * NLEAPS * NSTEPS            # leapfrog steps in one bench step, all ranks
    value = lf_per_step * args.steps / (dev_ms * 1e-3)
    out = None
    if rank == 0:
        # fp64-issue roofline (what binds this kernel, ncu: profiles/): executed leapfrog fp64 instructions, 5 d per
        # step in reference arithmetic (DADD / DMUL, un-fused) or 3 d (DFMA), one result per lane and instruction,
        # against the measured issue rate of a stream of independent DFMA on this device
        ops_launch = nloc * NSTEPS * NLEAPS * (5 if arith == "reference" else 3) * DIM
        fp64_achieved = ops_launch / (kernel_ms * 1e-3)
        # HBM contract figure of SURVEY.md 8d (read x + write x per transition, + the stored sample): kept beside it
        bytes_launch = nloc * (NSTEPS * 16 * DIM + NPOST * (8 * DIM + 9))
        hbm_achieved = bytes_launch / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = profiled_traffic()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": job_config(arith),
            "details": {"chains_per_gpu": nloc, "x0": "generated on the device (klb_job_set_state_synthetic)",
                       "closing_all_gather": ("none (N=1)" if world == 1 else
                                              "copy engines, CUDA IPC peer-to-peer over NVLink (klb_gather_*)" if gather is not None
                                              else "NCCL all_gather_into_tensor on a side stream"
                                              + (" (peer-to-peer setup failed: %s)" % gather_note if gather_note else "")),
                       "nv": plan.nv, "regs_per_thread": plan.regs_per_thread, "blocks_per_sm": plan.blocks_per_sm,
                       "accept_rate": acc_rate, "timing": "CUDA events on the job stream, max over ranks",
                       "wall_ms_per_step": wall_ms / args.steps},
            "roofline": {"bound": "fp64_issue", "achieved": fp64_achieved / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                         "frac": fp64_achieved / fp64_peak,
                         "flop_definition": "one fp64 result per lane and DADD / DMUL / DFMA instruction (an FMA counts 1): the pipe issues one "
                                            "warp instruction per 2 cycles per scheduler whatever the kind, so this is the issue-rate roofline",
                         "peak_source": "SMs x 64 fp64 lanes x clocks.max.sm (issue rate confirmed by microbenchmark, profiles/r1_summary.md); the "
                                        "live-measured stream of independent DFMA (klb_device_peak) is under `peaks`",
                         "traffic": None if not traffic else traffic["dram_bytes_per_launch"] * nloc / traffic["nchains"],
                         "traffic_source": traffic_src,
                         "kernel": "klb_hmc_ws_kernel<TgtIso, NV=16> (warp-specialised: 4 consumer + 4 producer warps per CTA)",
                         "kernel_ms": kernel_ms, "algorithmic_fp64_results_per_launch": ops_launch,
                         "hbm_contract": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                                          "algorithmic_bytes_per_launch": bytes_launch, "peak_source": peak_src,
                                          "note": "SURVEY.md 8d contract bytes (x read + written every transition, + stored samples). The kernel keeps "
                                                  "the state on chip, so measured DRAM traffic is ~5x lower and HBM is NOT the binding roofline"}},
            "cpu_baseline": cb,
            "e2e": {"value": lf_per_step / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms,
                    "call": "klb_job_set_state + klb_job_run + klb_job_output x3 (serial)" if args.e2e_serial
                            else "klb_job_run_host (chain slices pipelined over streams)",
                    "note": "host x0 in (pinned), final state + logtarget chain + accept flags out (pinned), through the "
                            "C ABI; the monitored values stay in HBM here -- e2e_full_output copies them too"},
            "e2e_full_output": None if e2e_full is None else (e2e_full if "unavailable" in e2e_full else dict(
                e2e_full, value=lf_per_step / (e2e_full["ms_per_step"] * 1e-3), unit=UNIT,
                note="as e2e, plus every monitored sample (KLB_OUT_VALUE, what output(job) holds in the reference) copied to "
                     "pinned host memory inside the same pipelined call: PCIe-bound")),
            "ess": {"mean_ess_per_coordinate": ess_sum / (NCHAINS * DIM), "min_ess": ess_min, "samples_per_chain": NPOST,
                    "independent_samples_per_sec": (ess_sum / DIM) / (dev_ms / args.steps * 1e-3),
                    "ess_kernel_ms": ess_ms,
                    "note": "ess(chain, :imse) per coordinate on the device (klb_job_ess); independent samples/s = "
                            "sum over chains of the coordinate-mean ESS / device time of one run"},
            "parity": parity,
            "configs": configs,
            "peaks": {"fp64_results_per_s": fp64_peak, "fp64_stream_measured_results_per_s": fp64_stream,
                      "dmma_flop_per_s": dmma_peak, "hbm_gbs": hbm_peak},
            "gpu_launches": int(launches_timed),
            "clocks": clocks,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


def verify(args, torch, K, L, job, nloc, lo, world, full_state, dev):
    """rank 0: the verification run against the CPU oracle (chains 0..n-1), and G-invariance probes at N > 1"""
    lib = L.lib()
    seconds = 12.0 if (world == 1 and not args.no_cpu) else 1.0
    cb, ref = cpu_leg(target_seconds=seconds)
    n = ref["accept"].shape[0]
    n = min(n, nloc)
    lp, _ = job.device_ptr(L.OUT_LOGTARGET)
    ap, _ = job.device_ptr(L.OUT_ACCEPT)
    vp, _ = job.device_ptr(L.OUT_VALUE)
    lt = _view(torch, lp, (nloc, NPOST), "f8", dev)[:n].cpu().numpy()
    ac = _view(torch, ap, (nloc, NPOST), "u1", dev)[:n].cpu().numpy()
    nv = min(n, 64, ref["value"].shape[0])
    val = _view(torch, vp, (nloc, NPOST, DIM), "f8", dev)[:nv].cpu().numpy()
    xf = full_state[:n].cpu().numpy()

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    def same(a, b):
        return bool(np.array_equal(a.view(np.uint64), b.view(np.uint64)))
    par = {"oracle_chains": int(n), "accept_equal": bool(np.array_equal(ac, ref["accept"][:n])),
           "max_rel_logtarget": rel(lt, ref["logtarget"][:n]), "max_rel_value": rel(val, ref["value"][:nv]),
           "max_rel_final_state": rel(xf, ref["x"][:n]),
           "bit_exact": same(lt, ref["logtarget"][:n]) and same(val, ref["value"][:nv]) and same(xf, ref["x"][:n]),
           "value_chains": int(nv)}
    if world == 1 and not args.no_cpu:
        par["_cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if world > 1:
        # G-invariance: the first chains of every OTHER rank's shard, recomputed here as their own small job
        probe, ok, checked = 32, True, 0
        p = K.BasicContMuvParameter("p", logtarget=K.IsoGaussian())
        for r in range(1, world):
            rlo, rhi = K.distributed.shard_range(NCHAINS, r, world)
            m = min(probe, rhi - rlo)
            pj = K.BasicMCJob(K.likelihood_model(p, False), K.HMC(LEAPSTEP, NLEAPS), K.BasicMCRange(nsteps=NSTEPS, burnin=BURNIN),
                              {"p": K.SyntheticNormal(m, DIM)}, outopts={"destination": "none"}, seed=SEED, arith=args.arith,
                              device=dev.index, chain_offset=rlo)
            pj.seek(VERIFY_T)
            pj.run()
            ok = ok and same(pj.pstate_value, full_state[rlo:rlo + m].cpu().numpy())
            checked += m
            pj.close()
        par["g_invariance"] = {"equal": bool(ok), "chains_recomputed_on_rank0": checked,
                               "what": "final states of chains owned by ranks 1..N-1 (as all-gathered) == the same chains "
                                       "run as a separate job on rank 0"}
    return par


def other_configs(args, torch, K, L, rank, world, local_rank, allmax, fp64_peak, dmma_peak, hbm_peak):
    """BASELINE.json configs C2, C4, C5, the MH sampler at the C3 size and HMC at dim 4096: device-resident runs, chains sharded over
    the ranks like C3, CUDA-event time of the run kernels (max over ranks), each with its own roofline figures."""
    idx = np.arange(512)
    Cm = np.linalg.inv(0.8 ** np.abs(idx[:, None] - idx[None, :]))
    Cm = np.ascontiguousarray((Cm + Cm.T) / 2)
    specs = [
        # name, workload, sampler, target, N, d, nsteps, burnin, tuner, unit, work units per transition
        ("C2", "MALA(0.9), -z.z, 4096 x 128, nsteps 2000 / burnin 1000", K.MALA(0.9), K.IsoGaussian(), 4096, 128, 2000, 1000, None, "transitions/s", 1),
        ("C4", "HMC(0.02, 20), -z'Cz (C = inv AR(1) 0.8, dense 512 x 512, DMMA), 16384 x 512, nsteps 200 / burnin 100",
         K.HMC(0.02, 20), K.DenseGaussian(Cm), 16384, 512, 200, 100, None, "leapfrog-steps/s", 20),
        ("C5", "MALA(0.01) + AcceptanceRateMCTuner(0.574), paired Rosenbrock, 32768 x 256, nsteps 2000 / burnin 1000",
         K.MALA(0.01), K.Rosenbrock(1.0, 100.0, 0.05), 32768, 256, 2000, 1000, K.AcceptanceRateMCTuner(0.574), "transitions/s", 1),
        ("MH", "MH(sigma = 0.02), -z.z, 65536 x 1024, nsteps 200 / burnin 100", K.MH(np.full(1024, 0.02)), K.IsoGaussian(),
         65536, 1024, 200, 100, None, "transitions/s", 1),
        ("HMC4096", "HMC(0.025, 10), -z.z, 16384 x 4096, nsteps 200 / burnin 100 (one chain per CTA: 4 consumer + 4 producer warps)",
         K.HMC(0.025, 10), K.IsoGaussian(), 16384, 4096, 200, 100, None, "leapfrog-steps/s", 10),
        ("NUTS", "NUTS(0.05; maxndoublings 5), -z.z, 65536 x 1024, nsteps 40 / burnin 20 (the reference's multivariate transition: 2^j leapfrog "
                 "steps per doubling, no U-turn stop)", K.NUTS(0.05), K.IsoGaussian(), 65536, 1024, 40, 20, None, "leapfrog-steps/s", 31),
    ]
    out = {}
    for name, workload, smp, tgt, N, d, nsteps, burnin, tuner, unit, per in specs:
        lo, hi = K.distributed.shard_range(N, rank, world)
        # A secondary configuration must never take the headline line down with it: its failure is recorded under its
        # name.  The collectives below run on every rank whatever happened, so the ranks stay in step.
        job, err, ms, acc, per_loc = None, None, [float("nan")], float("nan"), float(per)
        try:
            p = K.BasicContMuvParameter("p", logtarget=tgt)
            job = K.BasicMCJob(K.likelihood_model(p, False), smp, K.BasicMCRange(nsteps=nsteps, burnin=burnin),
                               {"p": K.SyntheticNormal(hi - lo, d)}, tuner=tuner,
                               outopts={"monitor": ["value", "logtarget"], "diagnostics": ["accept", "ndoublings"] if name == "NUTS" else ["accept"]},
                               seed=SEED, arith=args.arith, device=local_rank, chain_offset=lo)
            job.run()
            ms = []
            for _ in range(2):
                job.reset()
                job.run()
                ms.append(job.last_run_ms)
            acc = float(job.acceptance().mean())
            if name == "NUTS":                                                 # leapfrog steps per transition: 2^ndoublings - 1, from the diagnostic
                nd = job._fetch(L.OUT_NDOUBLINGS, (hi - lo, nsteps - burnin), np.uint8).astype(np.int64)
                per_loc = float((2 ** nd - 1).mean())
        except Exception as e:                                                 # noqa: BLE001  (recorded, not swallowed)
            err = "%s: %s" % (type(e).__name__, e)
        finally:
            if job is not None:
                try:
                    job.close()
                except Exception:                                              # noqa: BLE001
                    pass
        failed = allmax([0.0 if err is None else 1.0])[0] > 0
        per = float(allmax([per_loc])[0])
        ms = allmax([float(np.mean(ms)) if err is None else 0.0])[0]
        if failed:
            out[name] = {"workload": workload, "error": err or "failed on another rank"}
            continue
        nloc, npost = hi - lo, nsteps - burnin
        rate = N * nsteps * per / (ms * 1e-3)
        bytes_launch = nloc * (nsteps * 16 * d + npost * (8 * d + 9))          # SURVEY.md 8d contract bytes, this rank
        ent = {"workload": workload, "value": rate, "unit": unit, "ms_per_run": ms, "accept_rate": acc,
               "hbm_contract": {"achieved": bytes_launch / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": bytes_launch / (ms * 1e-3) / 1e9 / hbm_peak}}
        if name == "C4":
            flop = nloc * nsteps * per * (2 * d * d)                           # the matrix-vector products on the fp64 tensor pipe
            ent["roofline"] = {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12, "peak": dmma_peak / 1e12, "unit": "TFLOP/s",
                               "frac": flop / (ms * 1e-3) / dmma_peak, "kernel": "klb_dense_mma_kernel (DMMA m8n8k4 + cluster TMA multicast)",
                               "peak_source": "measured live: klb_device_peak(KLB_PEAK_DMMA), independent mma.sync.m8n8k4.f64"}
        elif name == "NUTS":
            res = nloc * nsteps * per * 8 * d                                  # 8 d fp64 operations per leaf (two half-kicks, the move, log-target, kinetic energy)
            ent["roofline"] = {"bound": "fp64_issue", "achieved": res / (ms * 1e-3) / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                               "frac": res / (ms * 1e-3) / fp64_peak, "kernel": "klb_nuts_kernel<TgtIso, 16, 1>",
                               "leapfrog_steps_per_transition": per}
        elif name == "HMC4096":
            res = nloc * nsteps * per * 5 * d                                  # the 5 d fp64 operations of a leapfrog step, as for C3
            ent["roofline"] = {"bound": "fp64_issue", "achieved": res / (ms * 1e-3) / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                               "frac": res / (ms * 1e-3) / fp64_peak, "kernel": "klb_hmc_ws_kernel<TgtIso, 16, W = 4>"}
        else:
            # algorithmic un-fused fp64 operations per transition: MALA iso 19 d, MALA Rosenbrock 25 d, MH 4 d
            ops = {"C2": 19, "C5": 25, "MH": 4}[name] * d
            res = nloc * nsteps * ops
            ent["roofline"] = {"bound": "issue (RNG + fp64)", "achieved": res / (ms * 1e-3) / 1e12, "peak": fp64_peak / 1e12,
                               "unit": "TFLOP/s", "frac": res / (ms * 1e-3) / fp64_peak,
                               "kernel": "klb_chain_kernel<%s, ...>" % type(smp).__name__,
                               "note": "algorithmic fp64 operations of SURVEY.md 8d against the measured fp64 issue rate; these kernels spend "
                                       "most of their issue slots on the d normals per transition (Philox + ziggurat) and, for MALA, on two fp64 "
                                       "divisions per element: ncu figures in profiles/"}
        out[name] = ent
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arith", default="reference", choices=["reference", "fma"])
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="closing all-gather at N > 1: copy engines over CUDA IPC peer-to-peer (default) or NCCL")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline timing (the parity check still runs a 1 s oracle sample)")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations (C2, C4, C5, MH)")
    ap.add_argument("--no-e2e-full", action="store_true", help="skip the end-to-end step that also copies every monitored sample to the host")
    ap.add_argument("--per-step-sync", action="store_true")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e leg through the three blocking calls (set_state, run, output) instead of klb_job_run_host")
    ap.add_argument("--e2e-slices", type=int, default=0, help="chain slices of the pipelined e2e call (0 = library default)")
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
