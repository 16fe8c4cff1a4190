// weld.cu — kernel 4: vertex weld (+ degenerate-triangle removal) on the GPU.
//
// Reference behaviour (/root/reference/src/meshify.c:45-106 unify_vertices, :113-168
// remove_degenerate_triangles, src/radixsort.c:277-348):
//   key_i = (float) |p_i - p_0|  (FP64 distance rounded to f32); stable sort by key; greedy scan in
//   key order: an unassigned vertex becomes a head; every later vertex whose f32 key difference to
//   the head is < 1e-5f and whose FP64 distance to the head is < 1e-5f is (re)assigned to it; the
//   merged vertex takes the coordinates of the highest original index of its cluster.  Then every
//   triangle with a repeated vertex or a Kahan/Heron area below FLT_EPSILON is dropped.
//
// Here vertices are already unique per grid edge (mc.cu), so only the CANDIDATES flagged by the
// emit pass (vertices within 2e-5 of a grid corner — the only way two different edge vertices can
// be closer than the tolerance) go through the reference procedure: keys, a radix sort of
// (key bits << 32 | original index), a split of the sorted list wherever consecutive keys are
// >= tol apart (the reference's window can never cross such a gap), and the reference's greedy
// loop run independently per segment (one thread each; segments are a handful of vertices).
// Non-candidates have no partner within tolerance, so the result equals the reference's on the
// full vertex set.  all_candidates = 1 runs the same code on every vertex (generic weld hook).
#include "common.cuh"

__device__ __forceinline__ double dist_rn(const double *a, const double *b) {  // dx(): sqrt(sqr+sqr+sqr), no FMA
  double x = __dsub_rn(a[0], b[0]), y = __dsub_rn(a[1], b[1]), z = __dsub_rn(a[2], b[2]);
  double s = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
  return __dsqrt_rn(s);
}

__global__ void __launch_bounds__(256) k_weld_keys(const double *__restrict__ verts, const uint32_t *__restrict__ cand,
                                                   unsigned n, const double *__restrict__ p0,
                                                   uint64_t *__restrict__ keys) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t vid = cand ? cand[i] : i;
  double o[3] = {p0[0], p0[1], p0[2]};
  float k = (float)dist_rn(o, verts + 3 * (size_t)vid);
  keys[i] = ((uint64_t)__float_as_uint(k) << 32) | (uint64_t)vid;  // keys are >= 0: bit order == float order
}

// one thread per sorted position; the thread at the start of a segment runs the reference's greedy
// loop over its segment.  head[i] = sorted position of the head vertex i ended up with.
__global__ void __launch_bounds__(128) k_weld_resolve(const double *__restrict__ verts, const uint64_t *__restrict__ keys,
                                                      unsigned n, uint32_t *__restrict__ head) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tol = 0.00001f;
  const double told = (double)tol;
  float ki = __uint_as_float((uint32_t)(keys[i] >> 32));
  if (i > 0) {
    float kp = __uint_as_float((uint32_t)(keys[i - 1] >> 32));
    if (__fsub_rn(ki, kp) < tol) return;  // not a segment start
  }
  unsigned end = i + 1;
  while (end < n) {
    float a = __uint_as_float((uint32_t)(keys[end] >> 32)), b = __uint_as_float((uint32_t)(keys[end - 1] >> 32));
    if (!(__fsub_rn(a, b) < tol)) break;
    end++;
  }
  for (unsigned q = i; q < end; q++) head[q] = 0xffffffffu;
  for (unsigned q = i; q < end; q++) {
    if (head[q] != 0xffffffffu) continue;  // already assigned
    float k0 = __uint_as_float((uint32_t)(keys[q] >> 32));
    const double *pq = verts + 3 * (size_t)(uint32_t)keys[q];
    for (unsigned j = q; j < end; j++) {
      float kj = __uint_as_float((uint32_t)(keys[j] >> 32));
      if (!(__fsub_rn(kj, k0) < tol)) break;
      if (dist_rn(pq, verts + 3 * (size_t)(uint32_t)keys[j]) < told) head[j] = q;
    }
  }
}

// representative of a cluster = its highest original index (src/meshify.c:99-100)
__global__ void __launch_bounds__(256) k_weld_rep(const uint64_t *__restrict__ keys, unsigned n,
                                                  const uint32_t *__restrict__ head, uint32_t *__restrict__ rep) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicMax(&rep[head[i]], (uint32_t)keys[i]);
}

// remap[vid] = representative; keepflag[vid] = 0 for merged-away vertices
__global__ void __launch_bounds__(256) k_weld_mark(const uint64_t *__restrict__ keys, unsigned n,
                                                   const uint32_t *__restrict__ head, const uint32_t *__restrict__ rep,
                                                   uint32_t *__restrict__ remap, uint32_t *__restrict__ keepflag,
                                                   unsigned int *__restrict__ n_removed) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  bool removed = false;
  if (i < n) {
    uint32_t vid = (uint32_t)keys[i];
    uint32_t r = rep[head[i]];
    if (r != vid) {
      remap[vid] = r;
      keepflag[vid] = 0;
      removed = true;
    }
  }
  unsigned m = __ballot_sync(0xffffffffu, removed);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_removed, (unsigned)__popc(m));
}

__global__ void __launch_bounds__(256) k_iota_ones(uint32_t *__restrict__ remap, uint32_t *__restrict__ keepflag, unsigned n) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { remap[i] = i; keepflag[i] = 1u; }
}

__global__ void __launch_bounds__(256) k_compact_verts(const double *__restrict__ vin, double *__restrict__ vout,
                                                       const uint32_t *__restrict__ remap,
                                                       const uint32_t *__restrict__ newidx, unsigned n) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (remap[i] != i) return;  // merged away
  size_t o = 3 * (size_t)newidx[i];
  vout[o] = vin[3 * (size_t)i]; vout[o + 1] = vin[3 * (size_t)i + 1]; vout[o + 2] = vin[3 * (size_t)i + 2];
}

// remap triangle indices (optional) and flag the degenerate ones (src/meshify.c:118-145)
__global__ void __launch_bounds__(256) k_tri_remap_degen(int *__restrict__ tris, unsigned nt,
                                                         const double *__restrict__ verts,
                                                         const uint32_t *__restrict__ remap,
                                                         const uint32_t *__restrict__ newidx,
                                                         uint32_t *__restrict__ keepflag) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  int a = tris[3 * (size_t)i], b = tris[3 * (size_t)i + 1], c = tris[3 * (size_t)i + 2];
  if (remap) {
    a = (int)newidx[remap[a]]; b = (int)newidx[remap[b]]; c = (int)newidx[remap[c]];
    tris[3 * (size_t)i] = a; tris[3 * (size_t)i + 1] = b; tris[3 * (size_t)i + 2] = c;
  }
  const double *pa = verts + 3 * (size_t)a, *pb = verts + 3 * (size_t)b, *pc = verts + 3 * (size_t)c;
  double l = dist_rn(pa, pb), m = dist_rn(pa, pc), n = dist_rn(pb, pc);
  double cc = fmin(fmin(l, m), n), aa = fmax(fmax(l, m), n);
  double bb = __dsub_rn(__dsub_rn(__dadd_rn(__dadd_rn(l, m), n), aa), cc);
  double amb = __dsub_rn(aa, bb);
  double t1 = __dsub_rn(cc, amb);
  uint32_t keep = 1;
  if (t1 <= 0.0) keep = 0;
  else {
    double prod = __dmul_rn(__dmul_rn(__dmul_rn(__dadd_rn(aa, __dadd_rn(bb, cc)), t1), __dadd_rn(cc, amb)),
                            __dadd_rn(aa, __dsub_rn(bb, cc)));
    double area4 = __dmul_rn(0.25, __dsqrt_rn(prod));
    if (area4 < (double)FLT_EPSILON) keep = 0;
  }
  keepflag[i] = keep;
}

__global__ void __launch_bounds__(256) k_compact_tris(const int *__restrict__ tin, int *__restrict__ tout,
                                                      const uint32_t *__restrict__ keepflag_scanned,
                                                      const uint32_t *__restrict__ total, unsigned nt) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt) return;
  uint32_t me = keepflag_scanned[i];
  uint32_t next = (i + 1 < nt) ? keepflag_scanned[i + 1] : *total;
  if (next == me) return;  // flag was 0
  size_t o = 3 * (size_t)me;
  tout[o] = tin[3 * (size_t)i]; tout[o + 1] = tin[3 * (size_t)i + 1]; tout[o + 2] = tin[3 * (size_t)i + 2];
}

int b2m_weld_run(b2m_ctx *ctx, b2m_mesh_dev *mesh, int all_candidates, int backend, b2m_result *res) {
  b2m_scalars *d_sc = b2m_ptr<b2m_scalars>(ctx, BUF_SCALARS);
  const unsigned nv = mesh->nv, nt = mesh->nt;
  unsigned ncand = all_candidates ? nv : mesh->ncand;
  double *verts = mesh->verts;
  int *tris = mesh->tris;
  unsigned nv_out = nv;
  const uint32_t *remap = nullptr, *newidx = nullptr;
  cudaEvent_t e0 = ctx->ev[2 * B2M_T_WELD], e1 = ctx->ev[2 * B2M_T_WELD + 1];
  cudaEvent_t e2 = ctx->ev[2 * B2M_T_DEGEN], e3 = ctx->ev[2 * B2M_T_DEGEN + 1];
  CU_TRY(cudaEventRecord(e0, ctx->stream));
  if (ncand >= 2) {
    B2M_TRY(b2m_reserve(ctx, BUF_SORTA, (size_t)ncand * 8));
    B2M_TRY(b2m_reserve(ctx, BUF_TMP0, (size_t)ncand * 4));
    B2M_TRY(b2m_reserve(ctx, BUF_TMP1, (size_t)ncand * 4));
    uint64_t *keys = b2m_ptr<uint64_t>(ctx, BUF_SORTA);
    uint32_t *head = b2m_ptr<uint32_t>(ctx, BUF_TMP0), *rep = b2m_ptr<uint32_t>(ctx, BUF_TMP1);
    // key origin: the reference's pts[0].  Lewiner: vertex 0.  Classic: first soup vertex (mc.cu).
    const double *p0 = (backend == B2M_BACKEND_CLASSIC && !all_candidates) ? d_sc->pts0 : verts;
    KT_LAUNCH(ctx, "weld_keys", k_weld_keys<<<b2m_cdiv(ncand, 256), 256, 0, ctx->stream>>>(verts, all_candidates ? nullptr : b2m_ptr<uint32_t>(ctx, BUF_CAND),
                                                                ncand, p0, keys));
    B2M_TRY(b2m_sort_u64(ctx, keys, ncand, 64));
    KT_LAUNCH(ctx, "weld_resolve", k_weld_resolve<<<b2m_cdiv(ncand, 128), 128, 0, ctx->stream>>>(verts, keys, ncand, head));
    CU_TRY(cudaMemsetAsync(rep, 0, (size_t)ncand * 4, ctx->stream));
    KT_LAUNCH(ctx, "weld_rep", k_weld_rep<<<b2m_cdiv(ncand, 256), 256, 0, ctx->stream>>>(keys, ncand, head, rep));
    B2M_TRY(b2m_reserve(ctx, BUF_REMAP, (size_t)nv * 4));
    B2M_TRY(b2m_reserve(ctx, BUF_FLAGS, (size_t)(nv > nt ? nv : nt) * 4 + 16));
    uint32_t *rm = b2m_ptr<uint32_t>(ctx, BUF_REMAP), *fl = b2m_ptr<uint32_t>(ctx, BUF_FLAGS);
    KT_LAUNCH(ctx, "iota_ones", k_iota_ones<<<b2m_cdiv(nv, 256), 256, 0, ctx->stream>>>(rm, fl, nv));
    CU_TRY(cudaMemsetAsync(&d_sc->n_removed, 0, 4, ctx->stream));
    KT_LAUNCH(ctx, "weld_mark", k_weld_mark<<<b2m_cdiv(ncand, 256), 256, 0, ctx->stream>>>(keys, ncand, head, rep, rm, fl, &d_sc->n_removed));
    CU_TRY(cudaGetLastError());
    B2M_TRY(b2m_fetch_scalars(ctx));
    unsigned nrem = ctx->h_scalars->n_removed;
    if (nrem > 0) {
      B2M_TRY(b2m_exclusive_scan_u32(ctx, fl, fl, nv, nullptr));
      nv_out = nv - nrem;
      B2M_TRY(b2m_reserve(ctx, BUF_VERTS2, (size_t)nv_out * 24));
      double *v2 = b2m_ptr<double>(ctx, BUF_VERTS2);
      KT_LAUNCH(ctx, "compact_verts", k_compact_verts<<<b2m_cdiv(nv, 256), 256, 0, ctx->stream>>>(verts, v2, rm, fl, nv));
      verts = v2;
      remap = rm;
      // newidx lives in BUF_FLAGS, which the triangle pass also needs for its keep flags: move it
      B2M_TRY(b2m_reserve(ctx, BUF_SORTB, (size_t)nv * 4));
      CU_TRY(cudaMemcpyAsync(ctx->buf[BUF_SORTB].p, fl, (size_t)nv * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      newidx = b2m_ptr<uint32_t>(ctx, BUF_SORTB);
    }
  }
  CU_TRY(cudaEventRecord(e1, ctx->stream));
  CU_TRY(cudaEventRecord(e2, ctx->stream));
  // degenerate triangles
  B2M_TRY(b2m_reserve(ctx, BUF_FLAGS, (size_t)(nv > nt ? nv : nt) * 4 + 16));
  uint32_t *tf = b2m_ptr<uint32_t>(ctx, BUF_FLAGS);
  KT_LAUNCH(ctx, "tri_remap_degen", k_tri_remap_degen<<<b2m_cdiv(nt, 256), 256, 0, ctx->stream>>>(tris, nt, verts, remap, newidx, tf));
  B2M_TRY(b2m_exclusive_scan_u32(ctx, tf, tf, nt, &d_sc->n_tri_kept));
  CU_TRY(cudaGetLastError());
  B2M_TRY(b2m_fetch_scalars(ctx));
  unsigned nt_out = ctx->h_scalars->n_tri_kept;
  if (nt_out != nt) {
    B2M_TRY(b2m_reserve(ctx, BUF_TRIS2, (size_t)nt_out * 12));
    int *t2 = b2m_ptr<int>(ctx, BUF_TRIS2);
    KT_LAUNCH(ctx, "compact_tris", k_compact_tris<<<b2m_cdiv(nt, 256), 256, 0, ctx->stream>>>(tris, t2, tf, &d_sc->n_tri_kept, nt));
    tris = t2;
  }
  CU_TRY(cudaEventRecord(e3, ctx->stream));
  CU_TRY(cudaGetLastError());
  mesh->verts = verts;
  mesh->tris = tris;
  res->nverts = (int)nv_out;
  res->ntris = (int)nt_out;
  res->nmerged = (int)(nv - nv_out);
  res->ndegenerate = (int)(nt - nt_out);
  res->d_verts = verts;
  res->d_tris = tris;
  return B2M_OK;
}
