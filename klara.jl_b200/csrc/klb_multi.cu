// klb_multi.cu -- chains sharded over GPUs (include/klara_b200.h, "multi-GPU").
//
// The reference's only multi-job facility is run(job::Vector{MCJob}) = map(run, job) (src/jobs/jobs.jl:212): N
// independent chains.  Device g of G owns the contiguous chain block [g N/G, (g+1) N/G) of the `dim x N` state matrix;
// RNG streams are keyed by the GLOBAL chain index (klb_config.chain_offset), so results do not depend on G; nothing is
// exchanged while sampling.  One closing all-gather puts every device's final states and per-chain tuner records into
// every device's gather buffers.  It is done by the COPY ENGINES over NVLink / NVSwitch -- one peer-to-peer
// cudaMemcpyAsync per (destination, array), no kernel -- so it takes no SM time from the next run's fp64-bound kernel
// (round 1 measured the SM-based ncclAllGather at ~6 % of the 8-GPU step when it overlapped that kernel).
//
// Two ways in:
//   klb_gather_*  one end per process (one process per GPU, e.g. under torchrun): buffers are shared between the
//                 processes as CUDA IPC memory handles, exchanged by the caller (any byte all-gather will do);
//   klb_multi_*   one process driving G devices: `ngpus` jobs + their gather ends, the SURVEY 8b `ngpus` semantics.
// Everything here sits on the public job ABI (klb_job_*): no access to job internals.
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/klara_b200.h"

int klb_set_error(int code, const char* msg);   // klb_api.cu: fills klb_last_error()

static int mfail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return klb_set_error(code, buf);
}
#define MCK(call)                                                                                       \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return mfail(e_ == cudaErrorMemoryAllocation ? KLB_ENOMEM : KLB_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// what a gather end carries per chain: final state column, log-target, tuner step, {accepted, proposed, totproposed}
static const int kFields[4] = {KLB_OUT_STATE, KLB_OUT_STATE_LOGTARGET, KLB_OUT_TUNE_STEP, KLB_OUT_TUNE_COUNTERS};

struct klb_gather {
  klb_job* job;
  int world, rank, device;
  long long ntotal, base, nloc, ld, dim;   // base: global index of chain 0 of the logical job
  long long off;                            // this rank's first chain within the logical job
  size_t per_chain[4];                      // bytes per chain of the four arrays
  char* mine;                               // this rank's buffer: [state | lt | step | counters], ntotal chains each
  size_t arr_off[4], bytes;
  std::vector<char*> peer;                  // base address of every rank's buffer as seen from this process
  std::vector<bool> opened;                 // peer[r] came from cudaIpcOpenMemHandle
  cudaStream_t copy;
  cudaEvent_t ran, pushed;
  bool connected;
};

struct GatherHandle {                       // KLB_GATHER_HANDLE_BYTES
  cudaIpcMemHandle_t mem;                   // 64 bytes
  long long ntotal, off, nloc, ld;
  int rank, device;
  char pad[KLB_GATHER_HANDLE_BYTES - 64 - 4 * 8 - 2 * 4];
};
static_assert(sizeof(GatherHandle) == KLB_GATHER_HANDLE_BYTES, "handle layout");

extern "C" {

int klb_gather_create(klb_job* job, int32_t world, int32_t rank, int64_t nchains_total, int64_t first_chain, klb_gather** out) {
  if (!job || !out || world < 1 || rank < 0 || rank >= world) return mfail(KLB_EINVAL, "bad argument");
  *out = nullptr;
  klb_config c;
  klb_plan p;
  { int rc = klb_job_config(job, &c); if (rc) return rc; }
  { int rc = klb_job_plan(job, &p); if (rc) return rc; }
  const long long off = c.chain_offset - first_chain;
  if (off < 0 || off + c.nchains > nchains_total)
    return mfail(KLB_EINVAL, "job covers chains [%lld, %lld) but the logical job is [%lld, %lld)", (long long)c.chain_offset,
                 (long long)(c.chain_offset + c.nchains), (long long)first_chain, (long long)(first_chain + nchains_total));
  klb_gather* g = new (std::nothrow) klb_gather();
  if (!g) return mfail(KLB_ENOMEM, "host allocation failed");
  g->job = job; g->world = world; g->rank = rank; g->device = c.device;
  g->ntotal = nchains_total; g->base = first_chain; g->nloc = c.nchains; g->off = off; g->ld = p.ld; g->dim = c.dim;
  g->per_chain[0] = (size_t)p.ld * 8; g->per_chain[1] = 8; g->per_chain[2] = 8; g->per_chain[3] = 24;
  size_t o = 0;
  for (int f = 0; f < 4; ++f) { g->arr_off[f] = o; o += ((size_t)nchains_total * g->per_chain[f] + 255) & ~(size_t)255; }
  g->bytes = o;
  g->mine = nullptr; g->copy = nullptr; g->ran = nullptr; g->pushed = nullptr; g->connected = false;
  g->peer.assign(world, nullptr); g->opened.assign(world, false);
#define GCK(call)                                                                                       \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      klb_gather_destroy(g);                                                                            \
      return mfail(e_ == cudaErrorMemoryAllocation ? KLB_ENOMEM : KLB_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    }                                                                                                   \
  } while (0)
  GCK(cudaSetDevice(g->device));
  GCK(cudaMalloc(&g->mine, g->bytes));            // cudaMalloc (not a pool): exportable as an IPC handle
  GCK(cudaMemset(g->mine, 0, g->bytes));
  GCK(cudaStreamCreateWithFlags(&g->copy, cudaStreamNonBlocking));
  GCK(cudaEventCreateWithFlags(&g->ran, cudaEventDisableTiming));
  GCK(cudaEventCreateWithFlags(&g->pushed, cudaEventDisableTiming));
#undef GCK
  g->peer[rank] = g->mine;
  if (world == 1) g->connected = true;
  *out = g;
  return KLB_OK;
}

int klb_gather_handle(klb_gather* g, void* handle_out) {
  if (!g || !handle_out) return mfail(KLB_EINVAL, "null argument");
  GatherHandle h;
  memset(&h, 0, sizeof h);
  MCK(cudaSetDevice(g->device));
  MCK(cudaIpcGetMemHandle(&h.mem, g->mine));
  h.ntotal = g->ntotal; h.off = g->off; h.nloc = g->nloc; h.ld = g->ld; h.rank = g->rank; h.device = g->device;
  memcpy(handle_out, &h, sizeof h);
  return KLB_OK;
}

int klb_gather_connect(klb_gather* g, const void* handles) {
  if (!g || !handles) return mfail(KLB_EINVAL, "null argument");
  MCK(cudaSetDevice(g->device));
  const GatherHandle* h = (const GatherHandle*)handles;
  for (int r = 0; r < g->world; ++r) {
    if (h[r].rank != r || h[r].ntotal != g->ntotal || h[r].ld != g->ld)
      return mfail(KLB_EINVAL, "handle %d does not belong to this all-gather (rank %d, %lld chains, ld %lld)", r, h[r].rank,
                   (long long)h[r].ntotal, (long long)h[r].ld);
    if (r == g->rank) continue;
    void* p = nullptr;
    MCK(cudaIpcOpenMemHandle(&p, h[r].mem, cudaIpcMemLazyEnablePeerAccess));
    g->peer[r] = (char*)p; g->opened[r] = true;
  }
  g->connected = true;
  return KLB_OK;
}

// in-process variant: the other ends live in this process (klb_multi); `ends` in rank order
static int gather_connect_local(klb_gather* g, klb_gather* const* ends) {
  MCK(cudaSetDevice(g->device));
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank) continue;
    if (ends[r]->device != g->device) {
      int can = 0;
      MCK(cudaDeviceCanAccessPeer(&can, g->device, ends[r]->device));
      if (can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(ends[r]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) MCK(e);
        cudaGetLastError();
      }                                           // without peer access cudaMemcpyPeerAsync stages through the host
    }
    g->peer[r] = ends[r]->mine;
  }
  g->connected = true;
  return KLB_OK;
}

// Step 1 (job stream, right behind the run): this rank's shard -> its own block of its own buffer, a local
// device-to-device copy (~20 us for 64 MiB) that doubles as the snapshot the transfers read, so the job's next run may
// overwrite the live state at once.  Step 2 (copy stream): that block -> the same block of every other rank's buffer,
// peer-to-peer through the copy engines, concurrently with whatever the job does next.
int klb_gather_push_async(klb_gather* g) {
  if (!g) return mfail(KLB_EINVAL, "null argument");
  if (!g->connected) return mfail(KLB_ESTATE, "klb_gather_connect first");
  MCK(cudaSetDevice(g->device));
  cudaStream_t js = (cudaStream_t)klb_job_stream(g->job);
  MCK(cudaStreamWaitEvent(js, g->pushed, 0));     // the previous round's transfers have read the snapshot
  for (int f = 0; f < 4; ++f) {
    void* src; int64_t nb;
    { int rc = klb_job_device_ptr(g->job, kFields[f], &src, &nb); if (rc) return rc; }
    MCK(cudaMemcpyAsync(g->mine + g->arr_off[f] + (size_t)g->off * g->per_chain[f], src, (size_t)g->nloc * g->per_chain[f],
                        cudaMemcpyDeviceToDevice, js));
  }
  MCK(cudaEventRecord(g->ran, js));
  MCK(cudaStreamWaitEvent(g->copy, g->ran, 0));
  for (int f = 0; f < 4; ++f) {
    const size_t at = g->arr_off[f] + (size_t)g->off * g->per_chain[f], bytes = (size_t)g->nloc * g->per_chain[f];
    for (int q = 1; q < g->world; ++q) {          // start with the next rank so that the ranks do not all hit rank 0 first
      const int r = (g->rank + q) % g->world;
      MCK(cudaMemcpyAsync(g->peer[r] + at, g->mine + at, bytes, cudaMemcpyDefault, g->copy));
    }
  }
  MCK(cudaEventRecord(g->pushed, g->copy));
  return KLB_OK;
}

// the job's stream waits for the transfers of the last push (so that an event recorded on it afterwards times them)
int klb_gather_join(klb_gather* g) {
  if (!g) return mfail(KLB_EINVAL, "null argument");
  MCK(cudaSetDevice(g->device));
  MCK(cudaStreamWaitEvent((cudaStream_t)klb_job_stream(g->job), g->pushed, 0));
  return KLB_OK;
}

int klb_gather_sync(klb_gather* g) {
  if (!g) return mfail(KLB_EINVAL, "null argument");
  MCK(cudaSetDevice(g->device));
  MCK(cudaStreamSynchronize(g->copy));
  return KLB_OK;
}

int klb_gather_device_ptr(klb_gather* g, int field, void** dev_ptr, int64_t* nbytes) {
  if (!g || !dev_ptr || !nbytes) return mfail(KLB_EINVAL, "null argument");
  for (int f = 0; f < 4; ++f)
    if (kFields[f] == field) {
      *dev_ptr = g->mine + g->arr_off[f];
      *nbytes = (int64_t)((size_t)g->ntotal * g->per_chain[f]);
      return KLB_OK;
    }
  return mfail(KLB_EINVAL, "field %d is not part of the closing all-gather", field);
}

int klb_gather_output(klb_gather* g, int field, void* host_dst, int64_t nbytes) {
  if (!g || !host_dst) return mfail(KLB_EINVAL, "null argument");
  void* p; int64_t nb;
  { int rc = klb_gather_device_ptr(g, field, &p, &nb); if (rc) return rc; }
  const size_t dense = field == KLB_OUT_STATE ? (size_t)g->ntotal * (size_t)g->dim * 8 : (size_t)nb;
  if ((size_t)nbytes != dense) return mfail(KLB_EINVAL, "field %d holds %zu bytes, caller passed %lld", field, dense, (long long)nbytes);
  MCK(cudaSetDevice(g->device));
  if (field == KLB_OUT_STATE && g->ld != g->dim)
    MCK(cudaMemcpy2DAsync(host_dst, (size_t)g->dim * 8, p, (size_t)g->ld * 8, (size_t)g->dim * 8, (size_t)g->ntotal,
                          cudaMemcpyDeviceToHost, g->copy));
  else
    MCK(cudaMemcpyAsync(host_dst, p, dense, cudaMemcpyDeviceToHost, g->copy));
  MCK(cudaStreamSynchronize(g->copy));
  return KLB_OK;
}

// unmap the other ranks' buffers.  An exporting process should keep its buffer allocated until the importers have closed
// it: one-process-per-GPU callers disconnect, meet at their own barrier, then destroy.
int klb_gather_disconnect(klb_gather* g) {
  if (!g) return mfail(KLB_EINVAL, "null argument");
  MCK(cudaSetDevice(g->device));
  if (g->copy) MCK(cudaStreamSynchronize(g->copy));
  for (int r = 0; r < g->world; ++r)
    if (g->opened[r] && g->peer[r]) { cudaIpcCloseMemHandle(g->peer[r]); g->peer[r] = nullptr; g->opened[r] = false; }
  g->connected = g->world == 1;
  return KLB_OK;
}

void klb_gather_destroy(klb_gather* g) {
  if (!g) return;
  klb_gather_disconnect(g);
  cudaSetDevice(g->device);
  cudaFree(g->mine);
  if (g->ran) cudaEventDestroy(g->ran);
  if (g->pushed) cudaEventDestroy(g->pushed);
  if (g->copy) cudaStreamDestroy(g->copy);
  cudaGetLastError();
  delete g;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- klb_multi
struct klb_multi {
  int ngpus;
  long long ntotal, dim, base;
  std::vector<klb_job*> job;
  std::vector<klb_gather*> end;
  std::vector<long long> lo, n;             // shard g = chains [lo, lo + n) of the logical job
};

// blocks differ by at most one chain (klara.jl_b200/distributed.py shard_range)
static void shard(long long N, int g, int G, long long* lo, long long* n) {
  const long long b = N / G, r = N % G;
  *lo = g * b + (g < r ? g : r);
  *n = b + (g < r ? 1 : 0);
}

extern "C" {

int klb_multi_create(const klb_config* cfg, int32_t ngpus, const int32_t* devices, klb_multi** out) {
  if (!cfg || !out) return mfail(KLB_EINVAL, "null argument");
  *out = nullptr;
  int ndev = klb_device_count();
  if (ndev == 0) return mfail(KLB_ECUDA, "no CUDA device available (this library has no CPU path)");
  if (ngpus <= 0) ngpus = ndev;
  if (cfg->nchains < ngpus) return mfail(KLB_EINVAL, "%lld chains cannot be sharded over %d devices", (long long)cfg->nchains, ngpus);
  klb_multi* m = new (std::nothrow) klb_multi();
  if (!m) return mfail(KLB_ENOMEM, "host allocation failed");
  m->ngpus = ngpus; m->ntotal = cfg->nchains; m->dim = cfg->dim; m->base = cfg->chain_offset;
  m->job.assign(ngpus, nullptr); m->end.assign(ngpus, nullptr); m->lo.assign(ngpus, 0); m->n.assign(ngpus, 0);
  for (int g = 0; g < ngpus; ++g) {
    klb_config c = *cfg;
    shard(cfg->nchains, g, ngpus, &m->lo[g], &m->n[g]);
    c.nchains = m->n[g];
    c.chain_offset = cfg->chain_offset + m->lo[g];
    c.device = devices ? devices[g] : g % ndev;     // the same device may appear twice (tests on a one-GPU box)
    int rc = klb_job_create(&c, &m->job[g]);
    if (rc == KLB_OK) rc = klb_gather_create(m->job[g], ngpus, g, cfg->nchains, cfg->chain_offset, &m->end[g]);
    if (rc) { klb_multi_destroy(m); return rc; }
  }
  for (int g = 0; g < ngpus; ++g) {
    const int rc = gather_connect_local(m->end[g], m->end.data());
    if (rc) { klb_multi_destroy(m); return rc; }
  }
  *out = m;
  return KLB_OK;
}

void klb_multi_destroy(klb_multi* m) {
  if (!m) return;
  for (klb_gather* e : m->end) klb_gather_destroy(e);
  for (klb_job* j : m->job) klb_job_destroy(j);
  delete m;
}

int klb_multi_ngpus(klb_multi* m) { return m ? m->ngpus : 0; }

int klb_multi_job(klb_multi* m, int32_t g, klb_job** job) {
  if (!m || !job || g < 0 || g >= m->ngpus) return mfail(KLB_EINVAL, "bad argument");
  *job = m->job[g];
  return KLB_OK;
}

int klb_multi_set_target_f64(klb_multi* m, int which, const double* host, int64_t n) {
  if (!m) return mfail(KLB_EINVAL, "null argument");
  for (klb_job* j : m->job) { int rc = klb_job_set_target_f64(j, which, host, n); if (rc) return rc; }
  return KLB_OK;
}

int klb_multi_set_state(klb_multi* m, const double* x0) {
  if (!m || !x0) return mfail(KLB_EINVAL, "null argument");
  for (int g = 0; g < m->ngpus; ++g) { int rc = klb_job_set_state(m->job[g], x0 + m->lo[g] * m->dim); if (rc) return rc; }
  return KLB_OK;
}

int klb_multi_set_state_synthetic(klb_multi* m) {
  if (!m) return mfail(KLB_EINVAL, "null argument");
  for (klb_job* j : m->job) { int rc = klb_job_set_state_synthetic(j); if (rc) return rc; }
  return KLB_OK;
}

int klb_multi_reset(klb_multi* m) {
  if (!m) return mfail(KLB_EINVAL, "null argument");
  for (klb_job* j : m->job) { int rc = klb_job_reset(j); if (rc) return rc; }
  return KLB_OK;
}

int klb_multi_seek(klb_multi* m, uint64_t t) {
  if (!m) return mfail(KLB_EINVAL, "null argument");
  for (klb_job* j : m->job) { int rc = klb_job_seek(j, t); if (rc) return rc; }
  return KLB_OK;
}

// run(job) on every device, then the closing all-gather; the launches of all devices are queued before anything waits
int klb_multi_run_async(klb_multi* m) {
  if (!m) return mfail(KLB_EINVAL, "null argument");
  for (klb_job* j : m->job) { int rc = klb_job_run_async(j); if (rc) return rc; }
  for (klb_gather* e : m->end) { int rc = klb_gather_push_async(e); if (rc) return rc; }
  return KLB_OK;
}

int klb_multi_sync(klb_multi* m) {
  if (!m) return mfail(KLB_EINVAL, "null argument");
  for (klb_job* j : m->job) { int rc = klb_job_sync(j); if (rc) return rc; }
  for (klb_gather* e : m->end) { int rc = klb_gather_sync(e); if (rc) return rc; }
  return KLB_OK;
}

int klb_multi_run(klb_multi* m) {
  const int rc = klb_multi_run_async(m);
  return rc ? rc : klb_multi_sync(m);
}

// reset(job, x0); run(job); output(job) from / to the host arrays of the logical job: every field is chain-major, so a
// shard's part of a host array starts at lo * (bytes per chain); all devices' pipelines are enqueued before any is awaited
int klb_multi_run_host(klb_multi* m, const double* x0, const klb_host_field* fields, int32_t nfields, int32_t nslices) {
  if (!m || (nfields > 0 && !fields) || nfields < 0 || nfields > 32) return mfail(KLB_EINVAL, "bad argument");
  int rc = KLB_OK, started = 0;
  for (int g = 0; g < m->ngpus && rc == KLB_OK; ++g) {
    klb_host_field f[32];
    for (int q = 0; q < nfields; ++q) {
      if (fields[q].nbytes % m->ntotal) return mfail(KLB_EINVAL, "field %d: %lld bytes is not a multiple of the chain count", fields[q].field, (long long)fields[q].nbytes);
      const int64_t per = fields[q].nbytes / m->ntotal;
      f[q] = fields[q];
      f[q].host_dst = (char*)fields[q].host_dst + m->lo[g] * per;
      f[q].nbytes = m->n[g] * per;
    }
    rc = klb_job_run_host_async(m->job[g], x0 ? x0 + m->lo[g] * m->dim : nullptr, f, nfields, nslices);
    if (rc == KLB_OK) ++started;
  }
  char keep[512] = "";
  if (rc != KLB_OK) snprintf(keep, sizeof keep, "%s", klb_last_error());
  for (int g = 0; g < started; ++g) {               // every started pipeline is awaited, whatever happened elsewhere
    const int r2 = klb_job_run_host_wait(m->job[g]);
    if (rc == KLB_OK && r2 != KLB_OK) { rc = r2; snprintf(keep, sizeof keep, "%s", klb_last_error()); }
  }
  // all or nothing: the shards of one logical job keep the same transition counter
  for (int g = 0; g < started; ++g) {
    if (rc == KLB_OK) klb_job_run_host_finish(m->job[g]);
    else klb_job_run_host_abort(m->job[g]);
  }
  if (rc != KLB_OK) return klb_set_error(rc, keep);
  for (klb_gather* e : m->end) { int r3 = klb_gather_push_async(e); if (r3) return r3; }
  for (klb_gather* e : m->end) { int r3 = klb_gather_sync(e); if (r3) return r3; }
  return KLB_OK;
}

// output(job): every field is chain-major, so the logical job's array is the concatenation of the shards
int klb_multi_output(klb_multi* m, int field, void* host_dst, int64_t nbytes) {
  if (!m || !host_dst) return mfail(KLB_EINVAL, "null argument");
  if (nbytes % m->ntotal) return mfail(KLB_EINVAL, "field %d: %lld bytes is not a multiple of the chain count", field, (long long)nbytes);
  const int64_t per = nbytes / m->ntotal;
  for (int g = 0; g < m->ngpus; ++g) {
    int rc = klb_job_output(m->job[g], field, (char*)host_dst + m->lo[g] * per, m->n[g] * per);
    if (rc) return rc;
  }
  return KLB_OK;
}

int klb_multi_gathered(klb_multi* m, int32_t g, int field, void** dev_ptr, int64_t* nbytes) {
  if (!m || g < 0 || g >= m->ngpus) return mfail(KLB_EINVAL, "bad argument");
  return klb_gather_device_ptr(m->end[g], field, dev_ptr, nbytes);
}

int klb_multi_gathered_output(klb_multi* m, int32_t g, int field, void* host_dst, int64_t nbytes) {
  if (!m || g < 0 || g >= m->ngpus) return mfail(KLB_EINVAL, "bad argument");
  return klb_gather_output(m->end[g], field, host_dst, nbytes);
}

double klb_multi_last_run_ms(klb_multi* m) {
  double worst = -1.0;
  if (m) for (klb_job* j : m->job) { const double ms = klb_job_last_run_ms(j); if (ms > worst) worst = ms; }
  return worst;
}

}  // extern "C"
