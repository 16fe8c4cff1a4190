// post.cu — SURVEY §8(f) rank 4: the post-smooth nii2mesh runs on the mesh after meshify() + apply_sform()
// (`-s <iterations>`, src/nii2mesh.c:330-335): Laplacian smoothing with Humphrey's classes.
//
// Reference behaviour reproduced bit-for-bit (/root/reference/src/quadric.c):
//   :315-341  laplacian_smooth(): ONE loop over the triangles scatters, per corner, the sum of the two other corners'
//             positions into sum[v] (FP64) and 2 into num[v]; new position = sum / num; vertices without a triangle stay
//   :343-369  laplacian_smoothHC(): q = p; p = L(p); b = p - (verts*alpha + q*(1-alpha)); q = L(b);
//             p = p - (b*beta + q*(1-beta)), iter times
//   :370-393  lockEdges: border vertices keep their original position; border rule of update_mesh(), :186-216 - a vertex
//             is a border vertex when the one-ring of SOME vertex (all corners of its incident triangles) holds it
//             exactly once
// Design: the triangle-order scatter of the reference is a serial dependency chain per vertex (FP64 addition does not
// associate), so the mesh gets a vertex -> (triangle, corner) incidence list once - counted and filled with atomics,
// then every vertex's few entries sorted by triangle - and each Laplacian is a GATHER, one thread per vertex, adding in
// exactly the reference's order.  The HC combinations are fused into the two gathers of an iteration.
#include "common.cuh"

#define PS_THREADS 256

__global__ void __launch_bounds__(PS_THREADS) k_ps_count(const int *__restrict__ tris, unsigned nt, unsigned nv, uint32_t *__restrict__ cnt,
                                                         unsigned int *__restrict__ bad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // corner index
  if (i >= 3 * (size_t)nt) return;
  const uint32_t v = (uint32_t)__ldg(tris + i);
  if (v >= nv) { atomicOr(bad, 1u); return; }
  atomicAdd(cnt + v, 1u);
}
__global__ void __launch_bounds__(PS_THREADS) k_ps_fill(const int *__restrict__ tris, unsigned nt, unsigned nv, const uint32_t *__restrict__ start,
                                                        uint32_t *__restrict__ fill, uint32_t *__restrict__ ref) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * (size_t)nt) return;
  const uint32_t v = (uint32_t)__ldg(tris + i);
  if (v >= nv) return;
  const uint32_t t = (uint32_t)(i / 3), c = (uint32_t)(i - 3 * (size_t)t);
  ref[start[v] + atomicAdd(fill + v, 1u)] = t * 4u + c;
}
// the atomics above fill a vertex's entries in any order: sort them (a handful per vertex) so that the gather adds in
// triangle order, then corner order - the order of the reference's loop
__global__ void __launch_bounds__(PS_THREADS) k_ps_sort(unsigned nv, const uint32_t *__restrict__ start, uint32_t *__restrict__ ref) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const uint32_t s = start[v], e = start[v + 1], n = e - s;
  uint32_t *a = ref + s;
  if (n <= 32u) {  // the usual handful: insertion sort
    for (uint32_t i = 1; i < n; i++) {
      const uint32_t key = a[i];
      uint32_t b = i;
      while (b > 0 && a[b - 1] > key) { a[b] = a[b - 1]; b--; }
      a[b] = key;
    }
    return;
  }
  // a vertex shared by very many triangles (the apex of a fan): heap sort keeps it O(n log n)
  auto sift = [&](uint32_t root, uint32_t end) {
    for (;;) {
      uint32_t c = 2 * root + 1;
      if (c >= end) return;
      if (c + 1 < end && a[c] < a[c + 1]) c++;
      if (a[root] >= a[c]) return;
      const uint32_t t = a[root]; a[root] = a[c]; a[c] = t;
      root = c;
    }
  };
  for (uint32_t i = n / 2; i-- > 0;) sift(i, n);
  for (uint32_t end = n - 1; end > 0; end--) {
    const uint32_t t = a[0]; a[0] = a[end]; a[end] = t;
    sift(0, end);
  }
}

struct ps_vec { double x, y, z; };
__device__ __forceinline__ ps_vec ps_ld(const double *__restrict__ a, size_t i) { return {a[3 * i], a[3 * i + 1], a[3 * i + 2]}; }
__device__ __forceinline__ void ps_st(double *__restrict__ a, size_t i, ps_vec v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }
// L(in)[v]  (src/quadric.c:315-341)
__device__ __forceinline__ ps_vec ps_laplacian(const double *__restrict__ in, const int *__restrict__ tris, const uint32_t *__restrict__ start,
                                               const uint32_t *__restrict__ ref, unsigned v) {
  double sx = 0.0, sy = 0.0, sz = 0.0;
  const uint32_t s = __ldg(start + v), e = __ldg(start + v + 1);
  for (uint32_t k = s; k < e; k++) {
    const uint32_t r = __ldg(ref + k);
    const int *t = tris + 3 * (size_t)(r >> 2);
    const uint32_t c = r & 3u;
    const int a = c == 0 ? __ldg(t + 1) : __ldg(t), b = c == 2 ? __ldg(t + 1) : __ldg(t + 2);
    const ps_vec pa = ps_ld(in, (size_t)a), pb = ps_ld(in, (size_t)b);
    sx = __dadd_rn(sx, __dadd_rn(pa.x, pb.x)); sy = __dadd_rn(sy, __dadd_rn(pa.y, pb.y)); sz = __dadd_rn(sz, __dadd_rn(pa.z, pb.z));
  }
  if (e == s) return ps_ld(in, v);
  const double num = (double)(int)(2u * (e - s));
  return {__ddiv_rn(sx, num), __ddiv_rn(sy, num), __ddiv_rn(sz, num)};
}
// first half of an HC iteration: p_new = L(p_old), b = p_new - (verts*alpha + p_old*(1-alpha))
__global__ void __launch_bounds__(PS_THREADS) k_ps_hc1(const double *__restrict__ verts, const double *__restrict__ p_old, double *__restrict__ p_new,
                                                       double *__restrict__ b, const int *__restrict__ tris, const uint32_t *__restrict__ start,
                                                       const uint32_t *__restrict__ ref, unsigned nv, double alpha, double alpha1) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const ps_vec l = ps_laplacian(p_old, tris, start, ref, v), o = ps_ld(verts, v), q = ps_ld(p_old, v);
  ps_st(p_new, v, l);
  ps_st(b, v, {__dsub_rn(l.x, __dadd_rn(__dmul_rn(o.x, alpha), __dmul_rn(q.x, alpha1))),
               __dsub_rn(l.y, __dadd_rn(__dmul_rn(o.y, alpha), __dmul_rn(q.y, alpha1))),
               __dsub_rn(l.z, __dadd_rn(__dmul_rn(o.z, alpha), __dmul_rn(q.z, alpha1)))});
}
// second half: p = p - (b*beta + L(b)*(1-beta)), in place on p
__global__ void __launch_bounds__(PS_THREADS) k_ps_hc2(double *__restrict__ p, const double *__restrict__ b, const int *__restrict__ tris,
                                                       const uint32_t *__restrict__ start, const uint32_t *__restrict__ ref, unsigned nv,
                                                       double beta, double beta1) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const ps_vec l = ps_laplacian(b, tris, start, ref, v), bb = ps_ld(b, v), pp = ps_ld(p, v);
  ps_st(p, v, {__dsub_rn(pp.x, __dadd_rn(__dmul_rn(bb.x, beta), __dmul_rn(l.x, beta1))),
               __dsub_rn(pp.y, __dadd_rn(__dmul_rn(bb.y, beta), __dmul_rn(l.y, beta1))),
               __dsub_rn(pp.z, __dadd_rn(__dmul_rn(bb.z, beta), __dmul_rn(l.z, beta1)))});
}
// border rule (src/quadric.c:186-216), one thread per vertex v: every id that occurs exactly once among the corners of
// v's incident triangles becomes a border vertex.  Quadratic in the one-ring size (18 corners for a regular vertex).
__global__ void __launch_bounds__(PS_THREADS) k_ps_border(const int *__restrict__ tris, const uint32_t *__restrict__ start,
                                                          const uint32_t *__restrict__ ref, unsigned nv, unsigned char *__restrict__ border) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const uint32_t s = __ldg(start + v), n = 3u * (__ldg(start + v + 1) - s);
  if (n <= 30u) {  // up to 10 incident triangles (a regular vertex has 6): the ids are read once
    int ids[30];
    for (uint32_t a = 0; a < n; a += 3) {
      const int *t = tris + 3 * (size_t)(__ldg(ref + s + a / 3u) >> 2);
      ids[a] = __ldg(t); ids[a + 1] = __ldg(t + 1); ids[a + 2] = __ldg(t + 2);
    }
    for (uint32_t a = 0; a < n; a++) {
      int cnt = 0;
      for (uint32_t c = 0; c < n; c++) cnt += ids[c] == ids[a];
      if (cnt == 1) border[ids[a]] = 1;
    }
    return;
  }
  for (uint32_t a = 0; a < n; a++) {
    const int ida = __ldg(tris + 3 * (size_t)(__ldg(ref + s + a / 3u) >> 2) + a % 3u);
    int cnt = 0;
    for (uint32_t c = 0; c < n; c++) cnt += __ldg(tris + 3 * (size_t)(__ldg(ref + s + c / 3u) >> 2) + c % 3u) == ida;
    if (cnt == 1) border[ida] = 1;
  }
}
__global__ void __launch_bounds__(PS_THREADS) k_ps_apply(double *__restrict__ verts, const double *__restrict__ p, const unsigned char *__restrict__ border,
                                                         unsigned nv) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  if (border && border[v]) return;
  ps_st(verts, v, ps_ld(p, v));
}

// laplacian_smoothHC() on a device-resident mesh (vec3d / vec3i layout), in place on d_verts
extern "C" int b2m_laplacian_hc_device(b2m_ctx *ctx, double *d_verts, const int *d_tris, int nvert, int ntri, double alpha,
                                       double beta, int iter, int lock_edges) {
  if (!ctx || nvert < 0 || ntri < 0 || (nvert && !d_verts) || (ntri && !d_tris)) { b2m_set_error("laplacian_hc: bad argument"); return B2M_EARG; }
  if ((unsigned long long)ntri * 4ull > 0xffffffffull) { b2m_set_error("laplacian_hc: more than 2^30 triangles"); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  if (nvert == 0) return B2M_OK;
  const unsigned nv = (unsigned)nvert, nt = (unsigned)ntri;
  const size_t nc = 3 * (size_t)nt, vb = (size_t)nv * 24;
  // workspace: [start nv+1 | fill nv | bad 1] u32, ref 3nt u32, p / p2 / b vertex arrays, border bytes
  B2M_TRY(b2m_reserve(ctx, BUF_POST_INC, ((size_t)2 * nv + 8) * 4 + (nc + 1) * 4));
  B2M_TRY(b2m_reserve(ctx, BUF_POST_P, 3 * vb + nv + 64));
  uint32_t *start = b2m_ptr<uint32_t>(ctx, BUF_POST_INC), *fill = start + nv + 1, *bad = fill + nv, *ref = bad + 7;
  double *pa = b2m_ptr<double>(ctx, BUF_POST_P), *pb = pa + 3 * (size_t)nv, *bv = pb + 3 * (size_t)nv;
  unsigned char *border = reinterpret_cast<unsigned char *>(bv + 3 * (size_t)nv);
  CU_TRY(cudaMemsetAsync(start, 0, ((size_t)2 * nv + 8) * 4, ctx->stream));
  const unsigned gc = b2m_cdiv(nc ? nc : 1, PS_THREADS), gv = b2m_cdiv(nv, PS_THREADS);
  if (nt) KT_LAUNCH(ctx, "ps_count", k_ps_count<<<gc, PS_THREADS, 0, ctx->stream>>>(d_tris, nt, nv, start, bad));
  unsigned hbad = 0;
  CU_TRY(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  B2M_TRY(b2m_exclusive_scan_u32(ctx, start, start, (size_t)nv + 1, nullptr));  // entry nv = total (its own count is 0)
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  if (hbad) { b2m_set_error("laplacian_hc: triangle index outside 0..nvert-1"); return B2M_EARG; }
  if (nt) {
    KT_LAUNCH(ctx, "ps_fill", k_ps_fill<<<gc, PS_THREADS, 0, ctx->stream>>>(d_tris, nt, nv, start, fill, ref));
    KT_LAUNCH(ctx, "ps_sort", k_ps_sort<<<gv, PS_THREADS, 0, ctx->stream>>>(nv, start, ref));
  }
  const double alpha1 = 1.0 - alpha, beta1 = 1.0 - beta;
  const double *cur = d_verts;  // p of the reference; ping-pongs between pa and pb
  for (int j = 0; j < iter; j++) {
    double *nxt = cur == pa ? pb : pa;
    KT_LAUNCH(ctx, "ps_hc1", k_ps_hc1<<<gv, PS_THREADS, 0, ctx->stream>>>(d_verts, cur, nxt, bv, d_tris, start, ref, nv, alpha, alpha1));
    KT_LAUNCH(ctx, "ps_hc2", k_ps_hc2<<<gv, PS_THREADS, 0, ctx->stream>>>(nxt, bv, d_tris, start, ref, nv, beta, beta1));
    cur = nxt;
  }
  if (lock_edges) {
    CU_TRY(cudaMemsetAsync(border, 0, nv, ctx->stream));
    KT_LAUNCH(ctx, "ps_border", k_ps_border<<<gv, PS_THREADS, 0, ctx->stream>>>(d_tris, start, ref, nv, border));
  }
  if (cur != d_verts) KT_LAUNCH(ctx, "ps_apply", k_ps_apply<<<gv, PS_THREADS, 0, ctx->stream>>>(d_verts, cur, lock_edges ? border : nullptr, nv));
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B2M_OK;
}

// the same from / to host memory (the reference's calling convention: verts updated in place)
extern "C" int b2m_laplacian_hc_host(b2m_ctx *ctx, double *h_verts, const int *h_tris, int nvert, int ntri, double alpha, double beta,
                                     int iter, int lock_edges) {
  if (!ctx || nvert < 0 || ntri < 0 || (nvert && !h_verts) || (ntri && !h_tris)) { b2m_set_error("laplacian_hc: bad argument"); return B2M_EARG; }
  if (nvert == 0) return B2M_OK;
  CU_TRY(cudaSetDevice(ctx->device));
  B2M_TRY(b2m_reserve(ctx, BUF_POST_V, (size_t)nvert * 24));
  B2M_TRY(b2m_reserve(ctx, BUF_POST_T, (size_t)(ntri ? ntri : 1) * 12));
  double *dv = b2m_ptr<double>(ctx, BUF_POST_V);
  int *dt = b2m_ptr<int>(ctx, BUF_POST_T);
  B2M_TRY(b2m_copy_h2d(ctx, dv, h_verts, (size_t)nvert * 24));
  if (ntri) B2M_TRY(b2m_copy_h2d(ctx, dt, h_tris, (size_t)ntri * 12));
  B2M_TRY(b2m_laplacian_hc_device(ctx, dv, dt, nvert, ntri, alpha, beta, iter, lock_edges));
  return b2m_copy_d2h(ctx, h_verts, dv, (size_t)nvert * 24);
}
