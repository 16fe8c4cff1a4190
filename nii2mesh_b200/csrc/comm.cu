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
#include <fcntl.h>
#include <nccl.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <atomic>

#include "common.cuh"

struct local_group {
  int world;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int waiting, generation, abort;
  const void *pub[64][2];  // per rank: published send pointers (up, down) or (send, -)
  int refs;
};

// Host decision points (b2m_sync_scalars) of ranks that share a node go through a POSIX shared-memory segment instead
// of an NCCL all-gather + D2H: every rank copies its own 256-byte block to the host and publishes it in the segment,
// then reads the other ranks' blocks - a few microseconds instead of the ~60-80 us of a small NCCL collective followed
// by a copy and a stream synchronisation, ~10 times per slab step.  Blocks are double-buffered by the parity of the
// sequence number (a rank cannot be two syncs ahead of a rank that is still reading).  The segment also carries the
// `poison` flag with which a rank that failed releases its peers (they abort the NCCL communicator and return an error
// instead of waiting for ever).
struct shm_slot {
  std::atomic<unsigned long long> seq;
  char pad[56];
  b2m_scalars blk[2];
};
struct shm_seg {
  std::atomic<unsigned int> poison;
  std::atomic<unsigned int> attached;
  std::atomic<unsigned int> magic;
  char pad[52];
  shm_slot r[1];  // [world]
};
#define SHM_MAGIC 0xb2b2b200u

struct b2m_comm {
  int rank, world, kind;  // kind 0 = NCCL, 1 = local
  ncclComm_t nccl;
  local_group *grp;
  shm_seg *seg;           // NCCL ranks of one node: host-side exchange of the scalar blocks (null: NCCL all-gather)
  size_t seg_bytes;
  unsigned long long seq;
  int aborted;
};

// ---- NCCL through dlopen -------------------------------------------------------------------------
static struct {
  void *h;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*CommAbort)(ncclComm_t);
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
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy") SYM(CommAbort, "ncclCommAbort")
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

static int shm_attach(b2m_comm *c, b2m_ctx *ctx, const void *id128);
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
// A rank that fails for a reason of its own (out of memory, a capacity overflow) must not leave its peers inside a
// collective.  Local groups: the abort flag of the barrier.  NCCL: the poison flag of the shared segment, which every
// host wait of the peers polls (b2m_comm_stream_wait, the sequence wait of b2m_sync_scalars), then ncclCommAbort so that
// operations already enqueued on the stream are released.  The communicator is unusable afterwards.
static void nccl_abort(b2m_comm *c) {
  if (c->aborted) return;
  c->aborted = 1;
  if (c->seg) c->seg->poison.store(1u, std::memory_order_release);
  // ncclCommAbort releases operations that are already enqueued; it is called from b2m_comm_destroy (an aborted
  // communicator is unusable either way), not here: this function runs inside host waits and must not block
}
void b2m_comm_abort(b2m_comm *c) {
  if (!c) return;
  if (c->kind == 0) { nccl_abort(c); return; }
  pthread_mutex_lock(&c->grp->mu);
  c->grp->abort = 1;
  pthread_cond_broadcast(&c->grp->cv);
  pthread_mutex_unlock(&c->grp->mu);
}


// ---- the node-local host segment of an NCCL communicator ----------------------------------------------
static unsigned long long fnv64(const void *p, size_t n, unsigned long long h) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
__global__ void k_comm_all_equal(const unsigned long long *v, int n, int *out) {
  int same = 1;
  for (int i = 1; i < n; i++) same &= v[i] == v[0];
  *out = same;
}
// All ranks decide together (one NCCL all-gather of a host fingerprint, once per communicator) whether they share a
// node; if so rank 0 creates the segment (name derived from the NCCL id), the others map it, and the name is removed
// again once everybody is attached.  B2M_SCALARS_NCCL=1 keeps the NCCL transport (tests run both).
static int shm_attach(b2m_comm *c, b2m_ctx *ctx, const void *id128) {
  const char *force = getenv("B2M_SCALARS_NCCL");
  char host[256] = "";
  gethostname(host, sizeof(host) - 1);
  unsigned long long fp = fnv64(host, strlen(host), 1469598103934665603ull);
  {
    char boot[64] = "";
    FILE *f = fopen("/proc/sys/kernel/random/boot_id", "r");
    if (f) { if (fgets(boot, sizeof(boot), f)) fp = fnv64(boot, strlen(boot), fp); fclose(f); }
  }
  if (force && atoi(force) > 0) fp = fnv64(&c->rank, sizeof(int), fp);  // pretend every rank sits on another node
  unsigned long long *d_fp = nullptr;
  int *d_same = nullptr, same = 0;
  CU_TRY(cudaMalloc(&d_fp, (size_t)(c->world + 1) * 8 + 8));
  d_same = reinterpret_cast<int *>(d_fp + c->world + 1);
  CU_TRY(cudaMemcpyAsync(d_fp + c->world, &fp, 8, cudaMemcpyHostToDevice, ctx->stream));
  NC_TRY(N.AllGather(d_fp + c->world, d_fp, 8, ncclUint8, c->nccl, ctx->stream));
  k_comm_all_equal<<<1, 1, 0, ctx->stream>>>(d_fp, c->world, d_same);
  CU_TRY(cudaMemcpyAsync(&same, d_same, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  CU_TRY(cudaFree(d_fp));
  if (!same) return B2M_OK;  // several nodes: the scalar blocks travel by NCCL
  char name[64];
  snprintf(name, sizeof(name), "/b2m_%016llx", fnv64(id128, 128, 1469598103934665603ull));
  const size_t bytes = sizeof(shm_seg) + (size_t)(c->world - 1) * sizeof(shm_slot);
  int fd = -1;
  if (c->rank == 0) {
    shm_unlink(name);
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { b2m_set_error("shm_open(%s) failed", name); if (fd >= 0) close(fd); return B2M_ECUDA; }
  } else {
    for (int tries = 0; tries < 60000 && fd < 0; tries++) {  // <= 60 s
      fd = shm_open(name, O_RDWR, 0600);
      struct stat st;
      if (fd >= 0 && (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes)) { close(fd); fd = -1; }
      if (fd < 0) usleep(1000);
    }
    if (fd < 0) { b2m_set_error("rank %d: cannot open the host segment %s", c->rank, name); return B2M_ECUDA; }
  }
  void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) { b2m_set_error("mmap of the host segment failed"); return B2M_ECUDA; }
  shm_seg *seg = (shm_seg *)m;
  if (c->rank == 0) seg->magic.store(SHM_MAGIC, std::memory_order_release);  // a fresh segment is zero-filled
  else
    for (int tries = 0; seg->magic.load(std::memory_order_acquire) != SHM_MAGIC; tries++) {
      if (tries > 60000) { munmap(m, bytes); b2m_set_error("host segment never initialised"); return B2M_ECUDA; }
      usleep(1000);
    }
  seg->attached.fetch_add(1u, std::memory_order_acq_rel);
  for (int tries = 0; seg->attached.load(std::memory_order_acquire) < (unsigned)c->world; tries++) {
    if (tries > 60000) { munmap(m, bytes); b2m_set_error("host segment: not every rank attached"); return B2M_ECUDA; }
    usleep(1000);
  }
  if (c->rank == 0) shm_unlink(name);
  c->seg = seg; c->seg_bytes = bytes;
  return B2M_OK;
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
  int rc = world > 1 ? shm_attach(c, ctx, id128) : B2M_OK;
  if (rc != B2M_OK) { N.CommDestroy(c->nccl); free(c); return rc; }
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
  if (c->kind == 0) {
    if (c->nccl) { if (c->aborted) N.CommAbort(c->nccl); else N.CommDestroy(c->nccl); }
    if (c->seg) munmap(c->seg, c->seg_bytes);
  } else {
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
// can the exchange run on a stream of its own, next to kernels of the ctx stream?  (NCCL: yes; local groups publish
// pointers between host barriers and always use the ctx stream)
bool b2m_comm_async_capable(const b2m_comm *c) { return c && c->world > 1 && c->kind == 0; }

int b2m_comm_exchange(b2m_ctx *ctx, b2m_comm *c, const void *d_send_up, size_t send_up_bytes, void *d_recv_lo,
                      size_t recv_lo_bytes, const void *d_send_dn, size_t send_dn_bytes, void *d_recv_hi,
                      size_t recv_hi_bytes) {
  return b2m_comm_exchange_on(ctx, c, ctx->stream, d_send_up, send_up_bytes, d_recv_lo, recv_lo_bytes, d_send_dn, send_dn_bytes,
                              d_recv_hi, recv_hi_bytes);
}
int b2m_comm_exchange_on(b2m_ctx *ctx, b2m_comm *c, cudaStream_t st, const void *d_send_up, size_t send_up_bytes, void *d_recv_lo,
                         size_t recv_lo_bytes, const void *d_send_dn, size_t send_dn_bytes, void *d_recv_hi,
                         size_t recv_hi_bytes) {
  if (!c || c->world == 1) return B2M_OK;
  const bool has_lo = c->rank > 0, has_hi = c->rank + 1 < c->world;
  if (c->kind == 0) {
    if (c->aborted) { b2m_set_error("nccl comm: aborted after a rank failed"); return B2M_ECUDA; }
    NC_TRY(N.GroupStart());
    if (has_hi && send_up_bytes) NC_TRY(N.Send(d_send_up, send_up_bytes, ncclUint8, c->rank + 1, c->nccl, st));
    if (has_lo && recv_lo_bytes) NC_TRY(N.Recv(d_recv_lo, recv_lo_bytes, ncclUint8, c->rank - 1, c->nccl, st));
    if (has_lo && send_dn_bytes) NC_TRY(N.Send(d_send_dn, send_dn_bytes, ncclUint8, c->rank - 1, c->nccl, st));
    if (has_hi && recv_hi_bytes) NC_TRY(N.Recv(d_recv_hi, recv_hi_bytes, ncclUint8, c->rank + 1, c->nccl, st));
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

// all list lengths are multiples of 8 bytes (pairs, entries, items)
struct gather_segs { int world; size_t stride; size_t off[65]; };
__global__ void __launch_bounds__(256) k_gather_pack(const char *__restrict__ scratch, char *__restrict__ out, gather_segs gs) {
  const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t o = u * 8;
  if (o >= gs.off[gs.world]) return;
  int r = 0;
  while (o >= gs.off[r + 1]) r++;
  *reinterpret_cast<unsigned long long *>(out + o) =
      *reinterpret_cast<const unsigned long long *>(scratch + (size_t)r * gs.stride + (o - gs.off[r]));
}
// fixed-size all-gather: every rank's `bytes` bytes at d_buf + rank * bytes, in place
int b2m_comm_allgather_inplace(b2m_ctx *ctx, b2m_comm *c, void *d_buf, size_t bytes) {
  const int W = b2m_comm_world(c), me = b2m_comm_rank(c);
  if (W == 1) return B2M_OK;
  if (c->kind == 0) {
    if (c->aborted) { b2m_set_error("nccl comm: aborted after a rank failed"); return B2M_ECUDA; }
    NC_TRY(N.AllGather((char *)d_buf + (size_t)me * bytes, d_buf, bytes, ncclUint8, c->nccl, ctx->stream));
    return B2M_OK;
  }
  local_group *g = c->grp;
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  g->pub[me][0] = (char *)d_buf + (size_t)me * bytes;
  B2M_TRY(grp_barrier(g));
  for (int r = 0; r < W; r++)
    if (r != me) CU_TRY(cudaMemcpyAsync((char *)d_buf + (size_t)r * bytes, g->pub[r][0], bytes, cudaMemcpyDefault, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  B2M_TRY(grp_barrier(g));
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
    // ONE ncclAllGather of equal-sized (padded to the longest list) segments into a scratch buffer, then one kernel
    // that packs the segments back to back (W grouped broadcasts were W times the latency)
    if (c->aborted) { b2m_set_error("nccl comm: aborted after a rank failed"); return B2M_ECUDA; }
    size_t mx = 0;
    for (int r = 0; r < W; r++) mx = bytes[r] > mx ? bytes[r] : mx;
    if (mx == 0) return B2M_OK;
    mx = (mx + 15) & ~(size_t)15;
    B2M_TRY(b2m_reserve(ctx, BUF_GATHER, mx * W));
    char *scratch = b2m_ptr<char>(ctx, BUF_GATHER);
    if (bytes[me]) CU_TRY(cudaMemcpyAsync(scratch + (size_t)me * mx, d_send, bytes[me], cudaMemcpyDeviceToDevice, ctx->stream));
    NC_TRY(N.AllGather(scratch + (size_t)me * mx, scratch, mx, ncclUint8, c->nccl, ctx->stream));
    gather_segs gs;
    gs.world = W; gs.stride = mx;
    for (int r = 0; r <= W; r++) gs.off[r] = off[r];
    const size_t units = (off[W] + 7) / 8;
    KT_LAUNCH(ctx, "gather_pack", k_gather_pack<<<b2m_cdiv(units, 256), 256, 0, ctx->stream>>>(scratch, (char *)d_recv, gs));
    CU_TRY(cudaGetLastError());
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

__global__ void k_comm_stamp(unsigned int *p, unsigned int v) { *p = v; }

// wait for the ctx stream from the host; in an NCCL group this polls the poison flag so that a failed peer cannot
// leave this rank inside a collective that will never complete
static double comm_now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static struct { int on; double stream_ms, peer_ms; unsigned long long n; } g_trace = {-1, 0, 0, 0};
int b2m_comm_stream_wait(b2m_ctx *ctx, b2m_comm *c) {
  static const bool poll = !(getenv("B2M_STREAM_POLL") && atoi(getenv("B2M_STREAM_POLL")) == 0);
  if (!c || c->kind != 0 || !c->seg || !poll) { CU_TRY(cudaStreamSynchronize(ctx->stream)); return B2M_OK; }
  for (unsigned spins = 0;; spins++) {
    const cudaError_t e = cudaStreamQuery(ctx->stream);
    if (e == cudaSuccess) return B2M_OK;
    if (e != cudaErrorNotReady) { b2m_set_error("stream: %s", cudaGetErrorString(e)); return B2M_ECUDA; }
    if (c->seg->poison.load(std::memory_order_acquire)) {
      nccl_abort(c);
      b2m_set_error("nccl comm: another rank failed");
      return B2M_ECUDA;
    }
    for (int k = 0; k < 200; k++) asm volatile("pause" ::: "memory");  // a few microseconds between two driver calls
  }
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
  if (c->kind == 0 && c->seg) {
    if (c->aborted) { b2m_set_error("nccl comm: aborted after a rank failed"); return B2M_ECUDA; }
    if (g_trace.on < 0) g_trace.on = getenv("B2M_SYNC_TRACE") && atoi(getenv("B2M_SYNC_TRACE")) > 0;
    const double tr0 = g_trace.on ? comm_now_ms() : 0.0;
    // The block is stamped with a sequence number on the device and copied into a pinned landing block; the host waits
    // by WATCHING that block (the stamp is its last word and PCIe writes arrive in order), not by asking the driver:
    // a thread that polls cudaStreamQuery holds the driver's locks most of the time and starves the threads NCCL needs
    // to progress the exchanges that are still in flight on the same stream (measured: 4 ms per sync instead of 20 us).
    static const bool watch = !(getenv("B2M_SYNC_WATCH") && atoi(getenv("B2M_SYNC_WATCH")) == 0);
    if (watch) {
      const unsigned stamp = ++ctx->sync_seq ? ctx->sync_seq : ++ctx->sync_seq;  // never 0
      k_comm_stamp<<<1, 1, 0, ctx->stream>>>(&reinterpret_cast<b2m_scalars *>(ctx->buf[BUF_SCALARS].p)->sync_seq, stamp);
      CU_TRY(cudaMemcpyAsync(ctx->h_land, mine, sizeof(b2m_scalars), cudaMemcpyDeviceToHost, ctx->stream));
      volatile unsigned *flag = &ctx->h_land->sync_seq;
      double t_start = 0.0;
      for (unsigned long long spins = 0; *flag != stamp; spins++) {
        asm volatile("pause" ::: "memory");
        if ((spins & 0xfffu) == 0xfffu) {
          if (c->seg->poison.load(std::memory_order_acquire)) { nccl_abort(c); b2m_set_error("nccl comm: another rank failed"); return B2M_ECUDA; }
          if ((spins & 0xfffffu) == 0xfffffu) {  // every ~1M spins: has the stream failed / is this taking minutes?
            const cudaError_t e = cudaStreamQuery(ctx->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) { b2m_set_error("stream: %s", cudaGetErrorString(e)); return B2M_ECUDA; }
            const double now = comm_now_ms();
            if (t_start == 0.0) t_start = now;
            if (now - t_start > 600e3) { nccl_abort(c); b2m_set_error("nccl comm: the stream did not reach the sync point in 10 minutes"); return B2M_ECUDA; }
          }
        }
      }
      __sync_synchronize();
      memcpy(ctx->h_scalars, ctx->h_land, sizeof(b2m_scalars));
    } else {
      CU_TRY(cudaMemcpyAsync(ctx->h_scalars, mine, sizeof(b2m_scalars), cudaMemcpyDeviceToHost, ctx->stream));
      B2M_TRY(b2m_comm_stream_wait(ctx, c));
    }
    const double tr1 = g_trace.on ? comm_now_ms() : 0.0;
    const unsigned long long s = ++c->seq;
    shm_slot *me = &c->seg->r[c->rank];
    memcpy(&me->blk[s & 1], ctx->h_scalars, sizeof(b2m_scalars));
    me->seq.store(s, std::memory_order_release);
    struct timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < W; r++) {
      shm_slot *o = &c->seg->r[r];
      for (unsigned spins = 0; o->seq.load(std::memory_order_acquire) < s; spins++) {
        if (spins < 4096) { asm volatile("pause" ::: "memory"); continue; }
        if (c->seg->poison.load(std::memory_order_acquire)) { nccl_abort(c); b2m_set_error("nccl comm: another rank failed"); return B2M_ECUDA; }
        sched_yield();
        if ((spins & 0xffffu) == 0) {
          struct timespec t1;
          clock_gettime(CLOCK_MONOTONIC, &t1);
          if (t1.tv_sec - t0.tv_sec > 180) { nccl_abort(c); b2m_set_error("nccl comm: rank %d did not reach sync %llu", r, s); return B2M_ECUDA; }
        }
      }
      memcpy(&ctx->h_all[r], &o->blk[s & 1], sizeof(b2m_scalars));
    }
    if (g_trace.on) {
      const double tr2 = comm_now_ms();
      g_trace.stream_ms += tr1 - tr0; g_trace.peer_ms += tr2 - tr1; g_trace.n++;
      if (g_trace.n % 64 == 0)
        fprintf(stderr, "[b2m sync trace] rank %d: %llu syncs, stream wait %.3f ms/sync, peer wait %.3f ms/sync\n", c->rank, g_trace.n,
                g_trace.stream_ms / g_trace.n, g_trace.peer_ms / g_trace.n);
    }
    return B2M_OK;
  }
  if (c->kind == 0) {
    if (c->aborted) { b2m_set_error("nccl comm: aborted after a rank failed"); return B2M_ECUDA; }
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
