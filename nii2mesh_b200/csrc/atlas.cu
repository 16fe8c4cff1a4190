// atlas.cu — the atlas front-end of nii2mesh on the GPU: one mesh per label of an indexed
// volume.
//
// Reference (/root/reference/src/nii2mesh.c:492-583): nLabel = trunc(max(img)); for every label i = 1..nLabel the WHOLE
// volume is binarised (`i-0.5 < img < i+0.5` -> 1 else 0, :553-563), labels without voxels are skipped (:564-567), and
// meshify() runs on the binary volume with isolevel 0.5 and -l forced off (:493, :571).  On the D99 atlas that is 365
// passes over 23 M voxels, almost all of them over zeros.
//
// Here: ONE pass over the label volume collects the voxel count and the bounding box of every label
// (b2m_atlas_scan); a label is then meshed on its bounding box grown by 4 voxels (clamped to the volume):
// b2m_meshify_label_device() crops + binarises that box and runs the ordinary pipeline on it, with the box origin
// added to the exported coordinates (in the reference's own single f32 / f64 addition, see mc.cu).  The result is the
// reference's mesh bit for bit, because zeros are invariant under every stage:
//   * the smooth spreads a label by 2 voxels per axis; the outermost TWO planes of a volume are passed through
//     unfiltered (meshify.c:187,198,208), so they must lie where the filtered value is 0 as well: distance >= 3,
//     hence the margin of 4;
//   * min = 0, max and the isolevel are those of the whole volume (a margin voxel is always inside the box, unless the
//     box is the whole volume);
//   * a face of the box that is not a face of the volume lies in the zero margin: darkening min(0.375, 0) changes
//     nothing there, and the background outside the box is one component that reaches the volume faces, so "touches a
//     box face" and "touches a volume face" select the same bubbles;
//   * the bright bounding box is widened by 1 / 2 voxels before marching cubes: inside the margin the clamps
//     max(lo-1, 0) / min(hi+2, dim) only bite where the box face IS the volume face.
#include <limits.h>

#include "common.cuh"

struct label_acc {  // device accumulator of one label
  unsigned long long n;
  int lo[3], hi[3];
};

// label of a voxel value: the i with i-0.5 < v < i+0.5 (float compares as the reference's), or 0
__device__ __forceinline__ int voxel_label(float v, int nlabel) {
  if (!(v > 0.5f)) return 0;
  const float r = rintf(v);
  if (!(r <= (float)nlabel)) return 0;
  const int i = (int)r;
  const float lo = (float)((double)i - 0.5), hi = (float)((double)i + 0.5);  // `float lo = i - 0.5` (nii2mesh.c:555-556)
  return (v > lo && v < hi) ? i : 0;
}

__global__ void __launch_bounds__(256) k_atlas_scan(const float *__restrict__ img, int nx, int ny, int nz, int nlabel,
                                                    label_acc *__restrict__ acc) {
  const long long rows = (long long)ny * nz;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int z = (int)(row / ny), y = (int)(row - (long long)z * ny);
    const float *r = img + row * nx;
    // a warp walks one row; lanes that hold the same label (the common case inside a region) combine first
    for (int x0 = 0; x0 < nx; x0 += 32) {
      const int x = x0 + (int)(threadIdx.x & 31);
      const int lab = x < nx ? voxel_label(__ldg(r + x), nlabel) : 0;
      const unsigned peers = __match_any_sync(0xffffffffu, lab);
      if (lab && (unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u)) {
        label_acc *a = acc + lab;
        atomicAdd(&a->n, (unsigned long long)__popc(peers));
        atomicMin(&a->lo[0], x0 + __ffs(peers) - 1);
        atomicMax(&a->hi[0], x0 + 31 - __clz(peers));
        atomicMin(&a->lo[1], y); atomicMax(&a->hi[1], y);
        atomicMin(&a->lo[2], z); atomicMax(&a->hi[2], z);
      }
    }
  }
}
__global__ void k_atlas_init(label_acc *acc, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  acc[i].n = 0;
  for (int a = 0; a < 3; a++) { acc[i].lo[a] = INT_MAX; acc[i].hi[a] = -1; }
}

// out = binary volume of `label` over the box [o, o+s)
__global__ void __launch_bounds__(256) k_atlas_crop(const float *__restrict__ img, int nx, int ny, int o0, int o1, int o2, int s0,
                                                    int s1, int s2, int label, float *__restrict__ out) {
  const size_t n = (size_t)s0 * s1 * s2;
  const float lo = (float)((double)label - 0.5), hi = (float)((double)label + 0.5);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % s0);
    const size_t t = i / s0;
    const int y = (int)(t % s1), z = (int)(t / s1);
    const float v = __ldg(img + ((size_t)(o2 + z) * ny + (o1 + y)) * nx + (o0 + x));
    out[i] = (v > lo && v < hi) ? 1.0f : 0.0f;
  }
}

extern "C" int b2m_atlas_scan(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], int *nlabel, b2m_label_info **infos) {
  if (!ctx || !d_img || !dims || !nlabel || !infos) { b2m_set_error("null argument"); return B2M_EARG; }
  for (int a = 0; a < 3; a++)
    if (dims[a] < 1 || dims[a] > 32767) { b2m_set_error("dims[%d] outside 1..32767", a); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  const size_t n = (size_t)dims[0] * dims[1] * dims[2];
  // max(img) -> nLabel = trunc(max) (nii2mesh.c:494-501)
  B2M_TRY(b2m_reserve(ctx, BUF_SCALARS, sizeof(b2m_scalars)));
  b2m_scalars *h = ctx->h_scalars;
  memset(h, 0, sizeof(*h));
  h->vmin_enc = 0xffffffffu;
  CU_TRY(cudaMemcpyAsync(ctx->buf[BUF_SCALARS].p, h, sizeof(*h), cudaMemcpyHostToDevice, ctx->stream));
  B2M_TRY(b2m_minmax_run(ctx, d_img, n, b2m_ptr<b2m_scalars>(ctx, BUF_SCALARS)));
  B2M_TRY(b2m_fetch_scalars(ctx));
  const float mx = f32_dec(h->vmax_enc);
  if (!(mx >= 1.0f)) {
    printf("intensity range not consistent with an indexed atlas %g..%g\n", f32_dec(h->vmin_enc), mx);
    return B2M_FAIL;
  }
  if (mx > 1.0e7f) { b2m_set_error("atlas: %g labels", mx); return B2M_EARG; }
  const int nl = (int)truncf(mx);
  B2M_TRY(b2m_reserve(ctx, BUF_ATLAS, (size_t)(nl + 1) * sizeof(label_acc)));
  label_acc *acc = b2m_ptr<label_acc>(ctx, BUF_ATLAS);
  KT_LAUNCH(ctx, "atlas_scan", k_atlas_init<<<b2m_cdiv(nl + 1, 256), 256, 0, ctx->stream>>>(acc, nl + 1));
  const long long rows = (long long)dims[1] * dims[2];
  const unsigned blocks = (unsigned)(rows / 8 + 1 < (long long)ctx->sm_count * 16 ? rows / 8 + 1 : (long long)ctx->sm_count * 16);
  KT_LAUNCH(ctx, "atlas_scan", k_atlas_scan<<<blocks, 256, 0, ctx->stream>>>(d_img, (int)dims[0], (int)dims[1], (int)dims[2], nl, acc));
  CU_TRY(cudaGetLastError());
  label_acc *hacc = (label_acc *)malloc((size_t)(nl + 1) * sizeof(label_acc));
  b2m_label_info *out = (b2m_label_info *)calloc((size_t)nl + 1, sizeof(b2m_label_info));
  if (!hacc || !out) { free(hacc); free(out); return B2M_ENOMEM; }
  cudaError_t e = cudaMemcpyAsync(hacc, acc, (size_t)(nl + 1) * sizeof(label_acc), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { free(hacc); free(out); b2m_set_error("atlas scan: %s", cudaGetErrorString(e)); return B2M_ECUDA; }
  for (int i = 0; i <= nl; i++) {
    out[i].label = i;
    out[i].nvox = (long long)hacc[i].n;
    for (int a = 0; a < 3; a++) { out[i].lo[a] = hacc[i].lo[a]; out[i].hi[a] = hacc[i].hi[a]; }
  }
  free(hacc);
  *nlabel = nl;
  *infos = out;
  return B2M_OK;
}

extern "C" void b2m_atlas_free(b2m_label_info *infos) { free(infos); }

#define ATLAS_MARGIN 4

extern "C" int b2m_meshify_label_device(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const b2m_label_info *info,
                                        const b2m_opts *opts, b2m_result *res) {
  if (!ctx || !d_img || !dims || !info || !opts || !res) { b2m_set_error("null argument"); return B2M_EARG; }
  if (info->nvox <= 0) { b2m_set_error("label %d has no voxels (the reference skips it)", info->label); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  int o[3], s[3];
  for (int a = 0; a < 3; a++) {
    if (info->lo[a] < 0 || info->hi[a] >= dims[a] || info->lo[a] > info->hi[a]) { b2m_set_error("bad label box"); return B2M_EARG; }
    o[a] = info->lo[a] - ATLAS_MARGIN > 0 ? info->lo[a] - ATLAS_MARGIN : 0;
    const int e = info->hi[a] + ATLAS_MARGIN < (int)dims[a] - 1 ? info->hi[a] + ATLAS_MARGIN : (int)dims[a] - 1;
    s[a] = e - o[a] + 1;
  }
  // the reference smooths only when EVERY dim of the volume is >= 5 (meshify.c:171): a box thinner than that inside a
  // larger volume would change the behaviour, so such (tiny) volumes are meshed whole
  for (int a = 0; a < 3; a++)
    if (s[a] < 5 && dims[a] >= 5) {
      o[a] = o[a] + 5 <= (int)dims[a] ? o[a] : (int)dims[a] - 5;
      s[a] = 5;
    }
  const size_t n = (size_t)s[0] * s[1] * s[2];
  B2M_TRY(b2m_reserve(ctx, BUF_INPUT, n * 4));
  float *sub = b2m_ptr<float>(ctx, BUF_INPUT);
  const unsigned blocks = (unsigned)((n + 255) / 256 < (size_t)ctx->sm_count * 32 ? (n + 255) / 256 : (size_t)ctx->sm_count * 32);
  k_atlas_crop<<<blocks, 256, 0, ctx->stream>>>(d_img, (int)dims[0], (int)dims[1], o[0], o[1], o[2], s[0], s[1], s[2], info->label, sub);
  CU_TRY(cudaGetLastError());
  b2m_opts lo = *opts;
  lo.only_largest = 0;  // nii2mesh.c:493
  const int64_t sd[3] = {s[0], s[1], s[2]};
  ctx->origin[0] = o[0]; ctx->origin[1] = o[1]; ctx->origin[2] = o[2];
  const int rc = b2m_meshify_device(ctx, sub, sd, &lo, res);
  ctx->origin[0] = ctx->origin[1] = ctx->origin[2] = 0;
  ctx->launches += 1;
  if (rc == B2M_OK) {
    res->launches += 1;
    for (int a = 0; a < 3; a++) { res->lo[a] += o[a]; res->hi[a] += o[a]; }  // bright box in the caller's coordinates
  }
  return rc;
}


// ---- all labels in one call ------------------------------------------------------------------------------------------
// The label loop of the reference (src/nii2mesh.c:540-579: one meshify() per non-empty label, under OpenMP when built
// with OMP=1) as ONE library call: a scan of the volume, then the labels - largest first - are handed out through an
// atomic counter to `workers` host threads that the library owns, each with its own b2m_ctx (stream + workspace) on the
// caller's device, so that the launch- and round-trip-bound per-label pipelines of different labels overlap on the GPU.
// The worker contexts live in the caller's ctx and are reused by later calls.
#include <pthread.h>
#include <atomic>
namespace {
struct atlas_job {
  const float *d_img;
  const int64_t *dims;
  const b2m_opts *opts;
  const b2m_label_info *infos;
  const int *order;
  int n;
  int fetch;
  b2m_label_mesh *out;
  std::atomic<int> next{0};
};
struct atlas_worker_arg { atlas_job *job; b2m_ctx *ctx; };
void *atlas_worker(void *a_) {
  atlas_worker_arg *a = (atlas_worker_arg *)a_;
  atlas_job *j = a->job;
  cudaSetDevice(a->ctx->device);
  for (;;) {
    const int k = j->next.fetch_add(1);
    if (k >= j->n) break;
    const int lab = j->order[k];
    b2m_label_mesh *m = &j->out[lab];
    m->rc = b2m_meshify_label_device(a->ctx, j->d_img, j->dims, &j->infos[lab], j->opts, &m->r);
    if (m->rc != B2M_OK) continue;
    m->nverts = m->r.nverts; m->ntris = m->r.ntris;
    if (j->fetch) {
      m->verts = malloc((size_t)m->nverts * 24 + 8);
      m->tris = malloc((size_t)m->ntris * 12 + 8);
      if (!m->verts || !m->tris) { m->rc = B2M_ENOMEM; continue; }
      m->rc = b2m_fetch_mesh(a->ctx, &m->r, m->verts, m->tris);
    }
    m->r.d_verts = nullptr; m->r.d_tris = nullptr;  // the worker's buffers are reused by its next label
  }
  return nullptr;
}
}  // namespace

extern "C" int b2m_atlas_meshify_all(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const b2m_opts *opts, int workers,
                                     int fetch, int *nlabel, b2m_label_mesh **meshes) {
  if (!ctx || !d_img || !dims || !opts || !nlabel || !meshes) { b2m_set_error("null argument"); return B2M_EARG; }
  if (workers < 1) workers = 8;
  if (workers > B2M_ATLAS_WORKERS_MAX) workers = B2M_ATLAS_WORKERS_MAX;
  b2m_label_info *infos = nullptr;
  int nl = 0;
  B2M_TRY(b2m_atlas_scan(ctx, d_img, dims, &nl, &infos));
  b2m_label_mesh *out = (b2m_label_mesh *)calloc((size_t)nl + 1, sizeof(b2m_label_mesh));
  int *order = (int *)malloc(((size_t)nl + 1) * sizeof(int));
  if (!out || !order) { free(out); free(order); b2m_atlas_free(infos); return B2M_ENOMEM; }
  int n = 0;
  for (int i = 0; i <= nl; i++) {
    out[i].label = i;
    out[i].nvox = infos[i].nvox;
    out[i].rc = -100;  // skipped: no voxels (src/nii2mesh.c:564-567)
    if (i >= 1 && infos[i].nvox > 0) order[n++] = i;
  }
  // largest boxes first: the long jobs start early and the small ones fill the gaps
  auto vol = [&](int l) { long long v = 1; for (int a = 0; a < 3; a++) v *= infos[l].hi[a] - infos[l].lo[a] + 9; return v; };
  for (int i = 1; i < n; i++) {
    const int l = order[i];
    const long long v = vol(l);
    int k = i - 1;
    while (k >= 0 && vol(order[k]) < v) { order[k + 1] = order[k]; k--; }
    order[k + 1] = l;
  }
  if (workers > n) workers = n > 0 ? n : 1;
  int rc = B2M_OK;
  for (int w = 0; w < workers && rc == B2M_OK; w++)
    if (!ctx->atlas_workers[w]) rc = b2m_create(&ctx->atlas_workers[w], ctx->device);
  if (rc == B2M_OK && n > 0) {
    atlas_job job;
    job.d_img = d_img; job.dims = dims; job.opts = opts; job.infos = infos; job.order = order; job.n = n; job.fetch = fetch; job.out = out;
    atlas_worker_arg args[B2M_ATLAS_WORKERS_MAX];
    pthread_t th[B2M_ATLAS_WORKERS_MAX];
    int started = 0;
    for (int w = 0; w < workers; w++) {
      args[w].job = &job; args[w].ctx = ctx->atlas_workers[w];
      if (pthread_create(&th[w], nullptr, atlas_worker, &args[w]) != 0) break;
      started++;
    }
    if (started == 0) { args[0].job = &job; args[0].ctx = ctx->atlas_workers[0]; atlas_worker(&args[0]); }
    for (int w = 0; w < started; w++) pthread_join(th[w], nullptr);
    CU_TRY(cudaSetDevice(ctx->device));
  }
  free(order);
  b2m_atlas_free(infos);
  if (rc != B2M_OK) { b2m_atlas_meshes_free(out, nl); return rc; }
  *nlabel = nl;
  *meshes = out;
  return B2M_OK;
}

extern "C" void b2m_atlas_meshes_free(b2m_label_mesh *m, int nlabel) {
  if (!m) return;
  for (int i = 0; i <= nlabel; i++) { free(m[i].verts); free(m[i].tris); }
  free(m);
}
