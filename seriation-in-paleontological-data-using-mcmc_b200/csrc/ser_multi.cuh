/* ser_multi.cuh -- the chains of one call sharded over the GPUs of a box (run_all_chains' Pool,
 * script.py:48-67, as one call) and the end-of-run exchange of the cross-chain step.
 * Part of the single translation unit ser_kernels.cu (included last).
 *
 * Two shapes:
 *   ser_comm_*   one process per GPU (torchrun / MPI): an NCCL communicator owned by the library; the
 *                all-gather of E[-logL] and the all-reduce of the pair-order counts run on the run's stream
 *                between the kernels of ser_run_cross_chain_async (no host synchronisation in between).
 *   ser_multi_*  one process, one host thread, n_gpus devices.  With peer access (NVLink / NVSwitch) the
 *                exchange is done by the kernels themselves: ser_stats_kernel stores its slice of E[-logL]
 *                into every device's gather buffer and ser_po_kernel stores the slab of a chosen chain into
 *                every device's count buffer (the slabs are disjoint: no reduction), ordered across devices
 *                by CUDA events.  Without peer access the same step runs over ncclCommInitAll communicators.
 * libnccl.so.2 is resolved at run time (the copy already loaded into the process -- e.g. PyTorch's -- or the
 * system one), so single-GPU users of the library do not need NCCL at all.
 */
#include <dlfcn.h>
#include <nccl.h>

/* ------------------------------------------------------------------ NCCL, resolved at run time */
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  const char *(*GetErrorString)(ncclResult_t);
  int ok;
};

static const NcclApi *nccl_api(void)
{
  static std::mutex mu;
  static NcclApi api;
  static int tried = 0;
  std::lock_guard<std::mutex> lock(mu);
  if (!tried) {
    tried = 1;
    void *h = nullptr;
    /* a copy already in the process first (so one NCCL serves PyTorch and this library), then the usual names */
    if (dlsym(RTLD_DEFAULT, "ncclCommInitRank")) h = RTLD_DEFAULT;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommInitAll = (decltype(api.CommInitAll))dlsym(h, "ncclCommInitAll");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather && api.AllReduce &&
               api.GroupStart && api.GroupEnd && api.GetErrorString;
    }
  }
  if (!api.ok) { ser_set_error("NCCL is not available (libnccl.so.2 could not be loaded): %s", dlerror() ? dlerror() : "missing symbols"); return nullptr; }
  return &api;
}

#define NCCL_TRY(api, expr)                                                                            \
  do {                                                                                                 \
    ncclResult_t r__ = (expr);                                                                         \
    if (r__ != ncclSuccess) {                                                                          \
      ser_set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, (api)->GetErrorString(r__));        \
      return SER_E_CUDA;                                                                               \
    }                                                                                                  \
  } while (0)

struct ser_comm {
  ncclComm_t comm;
  int n_ranks, rank, device;
};

static int comm_ranks(const ser_comm *comm, int *n_ranks, int *rank)
{
  if (!comm) return SER_E_ARG;
  *n_ranks = comm->n_ranks; *rank = comm->rank;
  return SER_OK;
}

/* in place: this rank's slice already sits at d_e_all + rank * n_local */
static int comm_all_gather_e(ser_comm *comm, double *d_e_all, int n_local, cudaStream_t stream)
{
  const NcclApi *api = nccl_api();
  if (!api) return SER_E_CUDA;
  NCCL_TRY(api, api->AllGather(d_e_all + (size_t)comm->rank * n_local, d_e_all, (size_t)n_local, ncclDouble, comm->comm, stream));
  return SER_OK;
}

static int comm_all_reduce_counts(ser_comm *comm, int *d_counts, size_t n, cudaStream_t stream)
{
  const NcclApi *api = nccl_api();
  if (!api) return SER_E_CUDA;
  NCCL_TRY(api, api->AllReduce(d_counts, d_counts, n, ncclInt32, ncclSum, comm->comm, stream));
  return SER_OK;
}

extern "C" int ser_comm_unique_id(uint8_t id[SER_COMM_ID_BYTES])
{
  static_assert(sizeof(ncclUniqueId) == SER_COMM_ID_BYTES, "ncclUniqueId size");
  if (!id) return SER_E_ARG;
  const NcclApi *api = nccl_api();
  if (!api) return SER_E_CUDA;
  ncclUniqueId u;
  NCCL_TRY(api, api->GetUniqueId(&u));
  memcpy(id, &u, SER_COMM_ID_BYTES);
  return SER_OK;
}

extern "C" int ser_comm_create(const uint8_t id[SER_COMM_ID_BYTES], int32_t n_ranks, int32_t rank, int32_t device, ser_comm **out)
{
  if (!id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) { ser_set_error("ser_comm_create: bad argument"); return SER_E_ARG; }
  *out = nullptr;
  const NcclApi *api = nccl_api();
  if (!api) return SER_E_CUDA;
  CUDA_TRY(cudaSetDevice(device));
  ser_comm *c = (ser_comm *)calloc(1, sizeof(ser_comm));
  if (!c) { ser_set_error("ser_comm_create: out of memory"); return SER_E_ARG; }
  ncclUniqueId u;
  memcpy(&u, id, SER_COMM_ID_BYTES);
  const ncclResult_t r = api->CommInitRank(&c->comm, n_ranks, u, rank);
  if (r != ncclSuccess) { ser_set_error("ncclCommInitRank: %s", api->GetErrorString(r)); free(c); return SER_E_CUDA; }
  c->n_ranks = n_ranks; c->rank = rank; c->device = device;
  *out = c;
  return SER_OK;
}

extern "C" int ser_comm_info(const ser_comm *comm, int32_t *n_ranks, int32_t *rank)
{
  if (!comm) return SER_E_ARG;
  if (n_ranks) *n_ranks = comm->n_ranks;
  if (rank) *rank = comm->rank;
  return SER_OK;
}

extern "C" void ser_comm_destroy(ser_comm *comm)
{
  if (!comm) return;
  const NcclApi *api = nccl_api();
  if (api && comm->comm) { cudaSetDevice(comm->device); api->CommDestroy(comm->comm); }
  free(comm);
}

/* ------------------------------------------------------------------ one process, n_gpus devices */
struct ser_multi {
  int n_gpus, n_total, first, k, N;
  int dev[SER_MAX_PEERS];
  int cnt[SER_MAX_PEERS], off[SER_MAX_PEERS + 1]; /* chains per device, first index of each device's block */
  ser_run *run[SER_MAX_PEERS];
  int peer;                          /* 1: peer stores; 0: NCCL */
  ncclComm_t comms[SER_MAX_PEERS];
  int have_comms;
  /* cross-chain buffers, one set per device (cudaMalloc: peer-mappable) */
  double *e_all[SER_MAX_PEERS], *info[SER_MAX_PEERS];
  int *chosen[SER_MAX_PEERS], *counts[SER_MAX_PEERS];
  cudaEvent_t ev_begin[SER_MAX_PEERS], ev_stats[SER_MAX_PEERS], ev_po[SER_MAX_PEERS];
};

static void multi_free_cc(ser_multi *m)
{
  for (int g = 0; g < m->n_gpus; g++) {
    cudaSetDevice(m->dev[g]);
    void *b[] = {m->e_all[g], m->info[g], m->chosen[g], m->counts[g]};
    for (void *p : b) if (p) cudaFree(p);
    m->e_all[g] = nullptr; m->info[g] = nullptr; m->chosen[g] = nullptr; m->counts[g] = nullptr;
  }
  m->k = 0;
}

extern "C" void ser_multi_destroy(ser_multi *m)
{
  if (!m) return;
  for (int g = 0; g < m->n_gpus; g++) if (m->run[g]) ser_run_sync(m->run[g]);
  multi_free_cc(m);
  if (m->have_comms) {
    const NcclApi *api = nccl_api();
    for (int g = 0; g < m->n_gpus; g++) if (api && m->comms[g]) { cudaSetDevice(m->dev[g]); api->CommDestroy(m->comms[g]); }
  }
  for (int g = 0; g < m->n_gpus; g++) {
    cudaSetDevice(m->dev[g]);
    if (m->ev_begin[g]) cudaEventDestroy(m->ev_begin[g]);
    if (m->ev_stats[g]) cudaEventDestroy(m->ev_stats[g]);
    if (m->ev_po[g]) cudaEventDestroy(m->ev_po[g]);
    ser_run_destroy(m->run[g]);
  }
  free(m);
}

static int multi_create_impl(ser_multi *m, const ser_dataset *ds, const ser_run_config *cfg)
{
  const int G = m->n_gpus;
  const int base = m->n_total / G, rem = m->n_total % G;
  m->off[0] = 0;
  for (int g = 0; g < G; g++) { m->cnt[g] = base + (g < rem); m->off[g + 1] = m->off[g] + m->cnt[g]; }
  /* peer access between every pair? */
  m->peer = 1;
  if (const char *v = getenv("SER_MULTI_NCCL")) if (atoi(v)) m->peer = 0;
  for (int g = 0; g < G && m->peer; g++)
    for (int h = 0; h < G && m->peer; h++) {
      if (g == h) continue;
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, m->dev[g], m->dev[h]));
      if (!can) m->peer = 0;
    }
  if (m->peer)
    for (int g = 0; g < G; g++) {
      CUDA_TRY(cudaSetDevice(m->dev[g]));
      for (int h = 0; h < G; h++) {
        if (g == h) continue;
        const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[h], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) { ser_set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", m->dev[g], m->dev[h], cudaGetErrorString(e)); return SER_E_CUDA; }
      }
    }
  if (!m->peer && G > 1) {
    if (rem) { ser_set_error("ser_multi_create: without peer access the chains must divide evenly over the devices (%d over %d)", m->n_total, G); return SER_E_ARG; }
    const NcclApi *api = nccl_api();
    if (!api) return SER_E_CUDA;
    NCCL_TRY(api, api->CommInitAll(m->comms, G, m->dev));
    m->have_comms = 1;
  }
  for (int g = 0; g < G; g++) {
    ser_run_config c = *cfg;
    c.n_chains = m->cnt[g];
    c.chain_offset = m->first + m->off[g];
    c.device = m->dev[g];
    const int rc = ser_run_create(ds, &c, &m->run[g]);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(m->dev[g]));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_begin[g], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_stats[g], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&m->ev_po[g], cudaEventDisableTiming));
  }
  return SER_OK;
}

extern "C" int ser_multi_create(const ser_dataset *ds, const ser_run_config *cfg, int32_t n_gpus, const int32_t *devices, ser_multi **out)
{
  if (!ds || !cfg || !out) { ser_set_error("ser_multi_create: null argument"); return SER_E_ARG; }
  *out = nullptr;
  if (cfg->struct_size != sizeof(ser_run_config)) { ser_set_error("ser_multi_create: cfg->struct_size does not match this library's ser_run_config"); return SER_E_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { ser_set_error("ser_multi_create: no CUDA device (this library has no CPU path)"); return SER_E_CUDA; }
  if (n_gpus < 1 || n_gpus > SER_MAX_PEERS || n_gpus > ndev) { ser_set_error("ser_multi_create: n_gpus=%d (devices present: %d, at most %d)", n_gpus, ndev, SER_MAX_PEERS); return SER_E_ARG; }
  if (cfg->n_chains < n_gpus) { ser_set_error("ser_multi_create: %d chains over %d devices", cfg->n_chains, n_gpus); return SER_E_ARG; }
  if (cfg->mode != SER_MODE_FREE) { ser_set_error("ser_multi_create: free-running chains only (replay tapes are per run: ser_run_set_tapes)"); return SER_E_ARG; }
  ser_multi *m = (ser_multi *)calloc(1, sizeof(ser_multi));
  if (!m) { ser_set_error("ser_multi_create: out of memory"); return SER_E_ARG; }
  m->n_gpus = n_gpus; m->n_total = cfg->n_chains; m->first = cfg->chain_offset; m->N = ds->N;
  for (int g = 0; g < n_gpus; g++) {
    m->dev[g] = devices ? devices[g] : g;
    if (m->dev[g] < 0 || m->dev[g] >= ndev) { ser_set_error("ser_multi_create: device %d of %d", m->dev[g], ndev); free(m); return SER_E_ARG; }
  }
  const int rc = multi_create_impl(m, ds, cfg);
  if (rc != SER_OK) {
    char keep[512];
    snprintf(keep, sizeof(keep), "%s", ser_last_error());
    ser_multi_destroy(m);
    cudaGetLastError();
    ser_set_error("%s", keep);
    return rc;
  }
  *out = m;
  return SER_OK;
}

extern "C" int ser_multi_init(ser_multi *m)
{
  if (!m) return SER_E_ARG;
  for (int g = 0; g < m->n_gpus; g++) { const int rc = ser_run_init(m->run[g]); if (rc) return rc; }
  return SER_OK;
}

extern "C" int ser_multi_advance(ser_multi *m, int32_t burn_calls, int32_t sample_calls)
{
  if (!m) return SER_E_ARG;
  for (int g = 0; g < m->n_gpus; g++) { const int rc = ser_run_advance_both(m->run[g], burn_calls, sample_calls); if (rc) return rc; }
  return SER_OK;
}

extern "C" int ser_multi_sync(ser_multi *m)
{
  if (!m) return SER_E_ARG;
  for (int g = 0; g < m->n_gpus; g++) { const int rc = ser_run_sync(m->run[g]); if (rc) return rc; }
  return SER_OK;
}

extern "C" int ser_multi_elapsed_ms(ser_multi *m, double *ms, int32_t reset)
{
  if (!m || !ms) return SER_E_ARG;
  double mx = 0.0;
  for (int g = 0; g < m->n_gpus; g++) {
    double t = 0.0;
    const int rc = ser_run_elapsed_ms(m->run[g], &t, reset);
    if (rc) return rc;
    if (t > mx) mx = t;
  }
  *ms = mx;
  return SER_OK;
}

extern "C" int ser_multi_layout(const ser_multi *m, int32_t *n_gpus, int32_t *chains_per_gpu, int32_t *uses_peer_stores)
{
  if (!m) return SER_E_ARG;
  if (n_gpus) *n_gpus = m->n_gpus;
  if (chains_per_gpu) for (int g = 0; g < m->n_gpus; g++) chains_per_gpu[g] = m->cnt[g];
  if (uses_peer_stores) *uses_peer_stores = m->peer;
  return SER_OK;
}

extern "C" int ser_multi_locate(ser_multi *m, int32_t global_chain, ser_run **run, int32_t *local_chain)
{
  if (!m) return SER_E_ARG;
  const int idx = global_chain - m->first;
  if (idx < 0 || idx >= m->n_total) { ser_set_error("ser_multi_locate: chain %d outside [%d, %d)", global_chain, m->first, m->first + m->n_total); return SER_E_ARG; }
  int g = 0;
  while (idx >= m->off[g + 1]) g++;
  if (run) *run = m->run[g];
  if (local_chain) *local_chain = idx - m->off[g];
  return SER_OK;
}

extern "C" int ser_multi_check(ser_multi *m, int32_t *n_bad)
{
  if (!m || !n_bad) return SER_E_ARG;
  int total = 0, worst = SER_OK;
  for (int g = 0; g < m->n_gpus; g++) {
    int32_t bad = 0;
    const int rc = ser_run_check(m->run[g], &bad);
    if (rc != SER_OK && rc != SER_E_CHECK) return rc;
    if (rc == SER_E_CHECK) worst = rc;
    total += bad;
  }
  *n_bad = total;
  if (worst) ser_set_error("ser_multi_check: %d inconsistent chain(s)", total);
  return worst;
}

extern "C" int ser_multi_chain_stats(ser_multi *m, double *e_negloglik, double *e_c, double *e_d, int32_t *n_samples)
{
  if (!m) return SER_E_ARG;
  for (int g = 0; g < m->n_gpus; g++) {
    const int o = m->off[g];
    const int rc = ser_run_chain_stats(m->run[g], e_negloglik ? e_negloglik + o : nullptr, e_c ? e_c + o : nullptr, e_d ? e_d + o : nullptr,
                                       g == 0 ? n_samples : nullptr);
    if (rc) return rc;
  }
  return SER_OK;
}

/* The cross-chain step over all devices, enqueued by one host thread without waiting for any device:
 *   barrier (events)  every stream waits until all devices have finished their earlier work
 *   stats             device g stores E[-logL] of its chains at [off_g, off_g + cnt_g) of EVERY device's gather buffer
 *   barrier (events)  ... so that after it every device holds all n_total values
 *   selection         every device runs the same single-CTA kernel on its own copy (script.py:70-99)
 *   pair order        the owner of a chosen chain stores the chain's N x N slab into EVERY device's count buffer
 *   device 0 waits for all pair-order events; the result is read from device 0. */
extern "C" int ser_multi_cross_chain(ser_multi *m, int32_t k, int32_t *chosen, int32_t *n_chosen, double *min_out, double *sigma_out,
                                     int32_t *counts)
{
  if (!m || k < 1) { ser_set_error("ser_multi_cross_chain: bad argument"); return SER_E_ARG; }
  const int G = m->n_gpus, N = m->N;
  for (int g = 0; g < G; g++)
    if (m->run[g]->cfg.store < SER_STORE_PI) { ser_set_error("ser_multi_cross_chain: runs have no pi sample store"); return SER_E_STATE; }
  if (m->k != k) {
    multi_free_cc(m);
    for (int g = 0; g < G; g++) {
      CUDA_TRY(cudaSetDevice(m->dev[g]));
      CUDA_TRY(cudaMalloc(&m->e_all[g], (size_t)m->n_total * sizeof(double)));
      CUDA_TRY(cudaMalloc(&m->info[g], 3 * sizeof(double)));
      CUDA_TRY(cudaMalloc(&m->chosen[g], (size_t)k * sizeof(int)));
      CUDA_TRY(cudaMalloc(&m->counts[g], (size_t)k * N * N * sizeof(int)));
    }
    m->k = k;
  }
  const size_t nn = (size_t)k * N * N;
  /* barrier: nobody writes into a peer that is still busy with earlier work */
  for (int g = 0; g < G; g++) { CUDA_TRY(cudaSetDevice(m->dev[g])); CUDA_TRY(cudaEventRecord(m->ev_begin[g], m->run[g]->stream)); }
  for (int g = 0; g < G; g++) {
    CUDA_TRY(cudaSetDevice(m->dev[g]));
    for (int h = 0; h < G; h++) if (h != g) CUDA_TRY(cudaStreamWaitEvent(m->run[g]->stream, m->ev_begin[h], 0));
  }
  const NcclApi *api = m->peer ? nullptr : nccl_api();
  if (!m->peer && G > 1 && !api) return SER_E_CUDA;
  /* stats (+ gather) */
  for (int g = 0; g < G; g++) {
    CUDA_TRY(cudaSetDevice(m->dev[g]));
    CUDA_TRY(cudaMemsetAsync(m->counts[g], 0, nn * sizeof(int), m->run[g]->stream));
    PeerPtrs dst;
    memset(&dst, 0, sizeof(dst));
    if (m->peer) { for (int h = 0; h < G; h++) dst.p[h] = m->e_all[h]; dst.n = G; }
    else { dst.p[0] = m->e_all[g]; dst.n = 1; }
    const int rc = run_stats_to(m->run[g], dst, m->off[g]);
    if (rc) return rc;
  }
  if (!m->peer && G > 1) {
    NCCL_TRY(api, api->GroupStart());
    for (int g = 0; g < G; g++)
      NCCL_TRY(api, api->AllGather(m->e_all[g] + m->off[g], m->e_all[g], (size_t)m->cnt[g], ncclDouble, m->comms[g], m->run[g]->stream));
    NCCL_TRY(api, api->GroupEnd());
  }
  for (int g = 0; g < G; g++) { CUDA_TRY(cudaSetDevice(m->dev[g])); CUDA_TRY(cudaEventRecord(m->ev_stats[g], m->run[g]->stream)); }
  /* selection + pair order */
  for (int g = 0; g < G; g++) {
    CUDA_TRY(cudaSetDevice(m->dev[g]));
    cudaStream_t st = m->run[g]->stream;
    if (m->peer) for (int h = 0; h < G; h++) if (h != g) CUDA_TRY(cudaStreamWaitEvent(st, m->ev_stats[h], 0));
    mark_launch(m->run[g]);
    ser_select_kernel<<<1, 1024, 0, st>>>(m->e_all[g], m->n_total, k, m->chosen[g], m->info[g]);
    CUDA_TRY(cudaGetLastError());
    PeerPtrs dst;
    memset(&dst, 0, sizeof(dst));
    if (m->peer) { for (int h = 0; h < G; h++) dst.p[h] = m->counts[h]; dst.n = G; }
    else { dst.p[0] = m->counts[g]; dst.n = 1; }
    const int rc = run_po_to(m->run[g], m->chosen[g], k, dst, m->off[g]);
    if (rc) return rc;
  }
  if (!m->peer && G > 1) {
    NCCL_TRY(api, api->GroupStart());
    for (int g = 0; g < G; g++) NCCL_TRY(api, api->AllReduce(m->counts[g], m->counts[g], nn, ncclInt32, ncclSum, m->comms[g], m->run[g]->stream));
    NCCL_TRY(api, api->GroupEnd());
  }
  for (int g = 0; g < G; g++) { CUDA_TRY(cudaSetDevice(m->dev[g])); CUDA_TRY(cudaEventRecord(m->ev_po[g], m->run[g]->stream)); }
  /* result from device 0 */
  CUDA_TRY(cudaSetDevice(m->dev[0]));
  cudaStream_t s0 = m->run[0]->stream;
  for (int h = 1; h < G; h++) CUDA_TRY(cudaStreamWaitEvent(s0, m->ev_po[h], 0));
  double info[3] = {0, 0, 0};
  std::vector<int> ch(k, -1);
  CUDA_TRY(cudaMemcpyAsync(info, m->info[0], sizeof(info), cudaMemcpyDeviceToHost, s0));
  CUDA_TRY(cudaMemcpyAsync(ch.data(), m->chosen[0], (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, s0));
  if (counts) CUDA_TRY(cudaMemcpyAsync(counts, m->counts[0], nn * sizeof(int), cudaMemcpyDeviceToHost, s0));
  CUDA_TRY(cudaStreamSynchronize(s0));
  for (int g = 1; g < G; g++) { const int rc = ser_run_sync(m->run[g]); if (rc) return rc; } /* the peers' buffers may be reused by the next call */
  if (chosen) for (int i = 0; i < k; i++) chosen[i] = ch[i] < 0 ? -1 : ch[i] + m->first;
  if (n_chosen) *n_chosen = (int)info[0];
  if (min_out) *min_out = info[1];
  if (sigma_out) *sigma_out = info[2];
  return SER_OK;
}
