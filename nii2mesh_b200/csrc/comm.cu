// comm.cu — the (small) communication layer of the z-slab path.
//
// One rank per GPU.  The data path needs four primitives, all on the ctx stream:
//   exchange     neighbour send/receive along z (halo planes, seam ids, seam vertices)
//   allgather    the 256-byte scalar blocks of all ranks (every host decision point)
//   allgatherv   variable-length lists (seam pairs / entries, weld items)
// Two back-ends behind one interface:
//   NCCL   one process per GPU (torchrun); libnccl.so.2 is dlopen()ed at run time so that the
//          single-GPU library has no NCCL dependency; ncclSend/ncclRecv groups over NVLink, ncclAllGather,
//          grouped ncclBroadcast for the variable-length lists.
//   LOCAL  `world` host threads of ONE process, each with its own b2m_ctx (same or different devices):
//          pointers are published through a shared table between two barriers and pulled with
//          cudaMemcpyAsync (peer copies when the devices differ).  This is what the parity tests use
//          to run 2..4 slabs on a single GPU, and it is a valid single-process multi-GPU mode.
#include <dlfcn.h>
#include <nccl.h>
#include <pthread.h>
#include <time.h>

#include "common.cuh"

struct local_group {
  int world;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int waiting, generation, abort;
  const void *pub[64][2];  // per rank: published send pointers (up, down) or (send, -)
  int refs;
};

struct b2m_comm {
  int rank, world, kind;  // kind 0 = NCCL, 1 = local
  ncclComm_t nccl;
  local_group *grp;
};

// ---- NCCL through dlopen -------------------------------------------------------------------------
static struct {
  void *h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
} N;

static int nccl_load(void) {
  if (N.h) return B2M_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  void *h = nullptr;
  for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { b2m_set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return B2M_ECUDA; }
#define SYM(field, name)                                                        \
  *(void **)(&N.field) = dlsym(h, name);                                        \
  if (!N.field) { b2m_set_error("libnccl: symbol %s missing", name); return B2M_ECUDA; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
  SYM(AllGather, "ncclAllGather") SYM(Broadcast, "ncclBroadcast") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  N.h = h;
  return B2M_OK;
}
#define NC_TRY(call)                                                                        \
  do {                                                                                      \
    ncclResult_t r_ = (call);                                                               \
    if (r_ != ncclSuccess) {                                                                \
      b2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, N.GetErrorString(r_));    \
      return B2M_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

int b2m_comm_rank(const b2m_comm *c) { return c ? c->rank : 0; }
int b2m_comm_world(const b2m_comm *c) { return c ? c->world : 1; }

// ---- local back-end: barrier with abort and timeout ---------------------------------------------
static int grp_barrier(local_group *g) {
  pthread_mutex_lock(&g->mu);
  if (g->abort) { pthread_mutex_unlock(&g->mu); b2m_set_error("local comm: another rank failed"); return B2M_ECUDA; }
  const int gen = g->generation;
  if (++g->waiting == g->world) {
    g->waiting = 0;
    g->generation++;
    pthread_cond_broadcast(&g->cv);
  } else {
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    ts.tv_sec += 120;
    while (gen == g->generation && !g->abort) {
      if (pthread_cond_timedwait(&g->cv, &g->mu, &ts) != 0) { g->abort = 1; pthread_cond_broadcast(&g->cv); break; }
    }
  }
  const int bad = g->abort;
  pthread_mutex_unlock(&g->mu);
  if (bad) { b2m_set_error("local comm: barrier aborted (another rank failed or timed out)"); return B2M_ECUDA; }
  return B2M_OK;
}
void b2m_comm_abort(b2m_comm *c) {
  if (!c || c->kind != 1) return;
  pthread_mutex_lock(&c->grp->mu);
  c->grp->abort = 1;
  pthread_cond_broadcast(&c->grp->cv);
  pthread_mutex_unlock(&c->grp->mu);
}

// ---- public constructors ---------------------------------------------------------------------------
extern "C" int b2m_comm_nccl_id(void *id128) {
  if (!id128) return B2M_EARG;
  B2M_TRY(nccl_load());
  ncclUniqueId id;
  NC_TRY(N.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return B2M_OK;
}
extern "C" int b2m_comm_create_nccl(b2m_comm **out, b2m_ctx *ctx, const void *id128, int rank, int world) {
  if (!out || !ctx || !id128 || world < 1 || world > 64 || rank < 0 || rank >= world) return B2M_EARG;
  B2M_TRY(nccl_load());
  CU_TRY(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  b2m_comm *c = (b2m_comm *)calloc(1, sizeof(b2m_comm));
  if (!c) return B2M_ENOMEM;
  c->rank = rank; c->world = world; c->kind = 0;
  ncclResult_t r = N.CommInitRank(&c->nccl, world, id, rank);
  if (r != ncclSuccess) { b2m_set_error("ncclCommInitRank: %s", N.GetErrorString(r)); free(c); return B2M_ECUDA; }
  *out = c;
  return B2M_OK;
}
extern "C" int b2m_comm_create_local(b2m_comm **out, int world) {
  if (!out || world < 1 || world > 64) return B2M_EARG;
  local_group *g = (local_group *)calloc(1, sizeof(local_group));
  if (!g) return B2M_ENOMEM;
  g->world = world; g->refs = world;
  pthread_mutex_init(&g->mu, nullptr);
  pthread_cond_init(&g->cv, nullptr);
  for (int r = 0; r < world; r++) {
    b2m_comm *c = (b2m_comm *)calloc(1, sizeof(b2m_comm));
    c->rank = r; c->world = world; c->kind = 1; c->grp = g;
    out[r] = c;
  }
  return B2M_OK;
}
extern "C" void b2m_comm_destroy(b2m_comm *c) {
  if (!c) return;
  if (c->kind == 0) { if (c->nccl) N.CommDestroy(c->nccl); }
  else {
    pthread_mutex_lock(&c->grp->mu);
    const int left = --c->grp->refs;
    pthread_mutex_unlock(&c->grp->mu);
    if (left == 0) { pthread_mutex_destroy(&c->grp->mu); pthread_cond_destroy(&c->grp->cv); free(c->grp); }
  }
  free(c);
}
extern "C" int b2m_comm_reset(b2m_comm *c) {  // clears the abort flag of a local group (tests that provoke failures)
  if (!c || c->kind != 1) return B2M_OK;
  pthread_mutex_lock(&c->grp->mu);
  c->grp->abort = 0; c->grp->waiting = 0;
  pthread_mutex_unlock(&c->grp->mu);
  return B2M_OK;
}

// ---- primitives --------------------------------------------------------------------------------------
int b2m_comm_exchange(b2m_ctx *ctx, b2m_comm *c, const void *d_send_up, size_t send_up_bytes, void *d_recv_lo,
                      size_t recv_lo_bytes, const void *d_send_dn, size_t send_dn_bytes, void *d_recv_hi,
                      size_t recv_hi_bytes) {
  if (!c || c->world == 1) return B2M_OK;
  const bool has_lo = c->rank > 0, has_hi = c->rank + 1 < c->world;
  if (c->kind == 0) {
    NC_TRY(N.GroupStart());
    if (has_hi && send_up_bytes) NC_TRY(N.Send(d_send_up, send_up_bytes, ncclUint8, c->rank + 1, c->nccl, ctx->stream));
    if (has_lo && recv_lo_bytes) NC_TRY(N.Recv(d_recv_lo, recv_lo_bytes, ncclUint8, c->rank - 1, c->nccl, ctx->stream));
    if (has_lo && send_dn_bytes) NC_TRY(N.Send(d_send_dn, send_dn_bytes, ncclUint8, c->rank - 1, c->nccl, ctx->stream));
    if (has_hi && recv_hi_bytes) NC_TRY(N.Recv(d_recv_hi, recv_hi_bytes, ncclUint8, c->rank + 1, c->nccl, ctx->stream));
    NC_TRY(N.GroupEnd());
    return B2M_OK;
  }
  local_group *g = c->grp;
  CU_TRY(cudaStreamSynchronize(ctx->stream));  // what we publish is complete
  g->pub[c->rank][0] = d_send_up;
  g->pub[c->rank][1] = d_send_dn;
  B2M_TRY(grp_barrier(g));
  if (has_lo && recv_lo_bytes) CU_TRY(cudaMemcpyAsync(d_recv_lo, g->pub[c->rank - 1][0], recv_lo_bytes, cudaMemcpyDefault, ctx->stream));
  if (has_hi && recv_hi_bytes) CU_TRY(cudaMemcpyAsync(d_recv_hi, g->pub[c->rank + 1][1], recv_hi_bytes, cudaMemcpyDefault, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  B2M_TRY(grp_barrier(g));  // nobody overwrites a published buffer before its reader is done
  return B2M_OK;
}

int b2m_comm_allgatherv(b2m_ctx *ctx, b2m_comm *c, const void *d_send, void *d_recv, const size_t *bytes) {
  const int W = b2m_comm_world(c), me = b2m_comm_rank(c);
  size_t off[65];
  off[0] = 0;
  for (int r = 0; r < W; r++) off[r + 1] = off[r] + bytes[r];
  if (W == 1) {
    if (bytes[0] && d_send != d_recv) CU_TRY(cudaMemcpyAsync(d_recv, d_send, bytes[0], cudaMemcpyDeviceToDevice, ctx->stream));
    return B2M_OK;
  }
  if (c->kind == 0) {
    NC_TRY(N.GroupStart());
    for (int r = 0; r < W; r++) {
      if (!bytes[r]) continue;
      NC_TRY(N.Broadcast(r == me ? d_send : nullptr, (char *)d_recv + off[r], bytes[r], ncclUint8, r, c->nccl, ctx->stream));
    }
    NC_TRY(N.GroupEnd());
    return B2M_OK;
  }
  local_group *g = c->grp;
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  g->pub[me][0] = d_send;
  B2M_TRY(grp_barrier(g));
  for (int r = 0; r < W; r++)
    if (bytes[r]) CU_TRY(cudaMemcpyAsync((char *)d_recv + off[r], g->pub[r][0], bytes[r], cudaMemcpyDefault, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  B2M_TRY(grp_barrier(g));
  return B2M_OK;
}

// the scalar blocks of all ranks -> pinned host (ctx->h_all[r]); also refreshes ctx->h_scalars
int b2m_sync_scalars(b2m_ctx *ctx, b2m_comm *c) {
  const int W = b2m_comm_world(c);
  if (W == 1) return b2m_fetch_scalars(ctx);
  if (ctx->h_all_cap < W) {
    if (ctx->h_all) { cudaFreeHost(ctx->h_all); cudaFree(ctx->d_all); ctx->h_all = nullptr; ctx->d_all = nullptr; }
    CU_TRY(cudaMallocHost((void **)&ctx->h_all, (size_t)W * sizeof(b2m_scalars)));
    CU_TRY(cudaMalloc((void **)&ctx->d_all, (size_t)W * sizeof(b2m_scalars)));
    ctx->h_all_cap = W;
  }
  const void *mine = ctx->buf[BUF_SCALARS].p;
  if (c->kind == 0) {
    NC_TRY(N.AllGather(mine, ctx->d_all, sizeof(b2m_scalars), ncclUint8, c->nccl, ctx->stream));
  } else {
    size_t bytes[64];
    for (int r = 0; r < W; r++) bytes[r] = sizeof(b2m_scalars);
    B2M_TRY(b2m_comm_allgatherv(ctx, c, mine, ctx->d_all, bytes));
  }
  CU_TRY(cudaMemcpyAsync(ctx->h_all, ctx->d_all, (size_t)W * sizeof(b2m_scalars), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(ctx->h_scalars, &ctx->h_all[c->rank], sizeof(b2m_scalars));
  return B2M_OK;
}
b2m_scalars *b2m_sc(b2m_ctx *ctx, b2m_comm *c, int rank) {
  if (b2m_comm_world(c) == 1) return ctx->h_scalars;
  return &ctx->h_all[rank];
}

// weld items of all ranks, concatenated in rank order (call after a b2m_sync_scalars that carried n_cand)
int b2m_comm_gather_items(b2m_ctx *ctx, b2m_comm *c, unsigned n_local, b2m_item **items, unsigned *n) {
  const int W = b2m_comm_world(c);
  if (W == 1) {
    *items = b2m_ptr<b2m_item>(ctx, BUF_CAND);
    *n = n_local;
    return B2M_OK;
  }
  size_t bytes[64];
  size_t tot = 0;
  for (int r = 0; r < W; r++) {
    bytes[r] = (size_t)b2m_sc(ctx, c, r)->n_cand * sizeof(b2m_item);
    tot += bytes[r];
  }
  if (bytes[c->rank] != (size_t)n_local * sizeof(b2m_item)) { b2m_set_error("gather_items: stale item counts"); return B2M_EARG; }
  B2M_TRY(b2m_reserve(ctx, BUF_ITEMS_ALL, tot));
  B2M_TRY(b2m_comm_allgatherv(ctx, c, ctx->buf[BUF_CAND].p, ctx->buf[BUF_ITEMS_ALL].p, bytes));
  *items = b2m_ptr<b2m_item>(ctx, BUF_ITEMS_ALL);
  *n = (unsigned)(tot / sizeof(b2m_item));
  return B2M_OK;
}
