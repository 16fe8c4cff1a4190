// pipeline.cu — orchestration of the hot path behind the C ABI (include/b2m.h):
// b2m_meshify_device() = the reference's meshify() (/root/reference/src/meshify.c:286-389) on a
// device-resident volume; b2m_meshify_host() adds the H2D / D2H copies; stage hooks for parity.
#include <limits.h>

#include "common.cuh"

extern "C" int b2m_get_default_backend(void);

#include <sys/mman.h>
#include <time.h>
static void hint_hugepages(void *p, size_t n) {
#ifdef MADV_HUGEPAGE
  const uintptr_t a = ((uintptr_t)p + 0x1fffff) & ~(uintptr_t)0x1fffff, e = ((uintptr_t)p + n) & ~(uintptr_t)0x1fffff;
  if (e > a) madvise((void *)a, e - a, MADV_HUGEPAGE);  // best effort; ignored when THP is off
#else
  (void)p; (void)n;
#endif
}
static double wall_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int stage_begin(b2m_ctx *ctx, int st) {
  CU_TRY(cudaEventRecord(ctx->ev[2 * st], ctx->stream));
  return B2M_OK;
}
static int stage_end(b2m_ctx *ctx, int st) {
  CU_TRY(cudaEventRecord(ctx->ev[2 * st + 1], ctx->stream));
  ctx->ev_mask |= 1u << st;
  return B2M_OK;
}

static int collect_times(b2m_ctx *ctx, b2m_result *res) {
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  for (int st = 0; st < B2M_NSTAGE; st++) {
    res->ms[st] = 0.f;
    if (ctx->ev_mask & (1u << st)) CU_TRY(cudaEventElapsedTime(&res->ms[st], ctx->ev[2 * st], ctx->ev[2 * st + 1]));
  }
  res->launches = ctx->launches;
  return B2M_OK;
}

static int check_dims(const int64_t dims[3]) {
  for (int a = 0; a < 3; a++)
    if (dims[a] < 1 || dims[a] > 32767) {
      b2m_set_error("dims[%d] = %lld outside 1..32767 (meshify.h uses short dim[3])", a, (long long)dims[a]);
      return B2M_EARG;
    }
  if ((long long)dims[0] * dims[1] * dims[2] > 0x7fffffffll) {
    b2m_set_error("more than 2^31-1 voxels: the meshify() API counts vertices/voxels in int");
    return B2M_EARG;
  }
  return B2M_OK;
}

static b2m_slab single_slab(int nz) {
  b2m_slab sl;
  sl.rank = 0; sl.world = 1; sl.gnz = nz; sl.z0 = 0; sl.nzl = nz; sl.hl = 0; sl.hh = 0; sl.ez0 = 0; sl.nze = nz;
  return sl;
}

static int init_scalars(b2m_ctx *ctx) {
  B2M_TRY(b2m_reserve(ctx, BUF_SCALARS, sizeof(b2m_scalars)));
  b2m_scalars *h = ctx->h_scalars;
  memset(h, 0, sizeof(*h));
  h->vmin_enc = 0xffffffffu; h->vmax_enc = 0u; h->cmin_enc = 0xffffffffu;
  for (int a = 0; a < 3; a++) { h->lo[a] = INT_MAX; h->hi[a] = -1; }
  h->first_cube = ~0ull;
  CU_TRY(cudaMemcpyAsync(ctx->buf[BUF_SCALARS].p, h, sizeof(*h), cudaMemcpyHostToDevice, ctx->stream));
  return B2M_OK;
}

// smooth -> range -> isolevel sanity -> CC masks -> bright bbox.  src/meshify.c:299-371
// d_img = the own planes of this rank (nzl of them); g = geometry of the EXT planes (own + halo).
int b2m_front_run(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const float *d_img, const b2m_geom &g, const b2m_opts *o,
                  b2m_front_out *fo, b2m_result *res) {
  B2M_TRY(init_scalars(ctx));
  b2m_scalars *d_sc = b2m_ptr<b2m_scalars>(ctx, BUF_SCALARS);
  const bool slabs = sl.world > 1;
  const size_t nxy = (size_t)g.nxy;

  const bool smooth = o->pre_smooth && g.nx >= 5 && g.ny >= 5 && sl.gnz >= 5;  // meshify.c:171
  if (smooth) {
    B2M_TRY(stage_begin(ctx, B2M_T_SMOOTH));
    B2M_TRY(b2m_reserve(ctx, BUF_SMOOTH, (size_t)g.n * 4));
    float *S = b2m_ptr<float>(ctx, BUF_SMOOTH);
    smooth_src src;
    memset(&src, 0, sizeof(src));
    src.main = d_img; src.n_main = sl.nzl; src.rz0 = sl.z0; src.gnz = sl.gnz;
    src.lo = src.hi = d_img;
    src.oz0 = sl.ez0; src.onz = sl.nze;
    // threshold bit rows as a by-product of the smooth, for the isolevel the caller asked for.  Opt-in (B2M_SMOOTH_BITS=1):
    // bit-exact (tests/test_gpu_parity.py::test_smooth_writes_the_threshold_bit_rows runs it), but on B200 at G1024 the
    // smooth grows by 1.6-1.8 ms (it is issue-bound at 128 registers per thread) while k_threshold only costs 0.9 ms
    const char *eb = getenv("B2M_SMOOTH_BITS");  // read per call: the test switches it inside one process
    const bool want_bits = eb && atoi(eb) > 0;
    smooth_bits sb;
    memset(&sb, 0, sizeof(sb));
    const bool try_bits = want_bits && g.nx % 32 == 0 && o->isolevel == o->isolevel;
    const size_t pwords = (size_t)g.ny * g.w;
    if (try_bits) {
      B2M_TRY(b2m_reserve(ctx, BUF_FG, (size_t)g.nwords * 4));
      B2M_TRY(b2m_reserve(ctx, BUF_MB, (size_t)g.nwords * 4));
      if (o->fill_bubbles) B2M_TRY(b2m_reserve(ctx, BUF_BG, (size_t)g.nwords * 4));
      sb.fg = b2m_ptr<uint32_t>(ctx, BUF_FG);
      sb.mb = b2m_ptr<uint32_t>(ctx, BUF_MB);
      sb.bg = o->fill_bubbles ? b2m_ptr<uint32_t>(ctx, BUF_BG) : nullptr;
      sb.iso = o->isolevel;
      sb.classic = o->backend == B2M_BACKEND_CLASSIC;
    }
    int bits_all = try_bits ? 1 : 0;
    // one launch of the smooth over the output planes src.oz0 .. (first of them = plane `p0` of S)
    auto smooth_planes = [&](const smooth_src &sv, size_t p0) -> int {
      if (!try_bits) return b2m_smooth_run(ctx, sv, S + p0 * nxy, g, d_sc);
      smooth_bits b = sb;
      b.fg += p0 * pwords; b.mb += p0 * pwords;
      if (b.bg) b.bg += p0 * pwords;
      int done = 0;
      const int rc = b2m_smooth_run(ctx, sv, S + p0 * nxy, g, d_sc, &b, &done);
      bits_all &= done;
      return rc;
    };
    bool halo_async = false;
    if (slabs) {
      // 3 raw planes from each neighbour: the halo plane of S is recomputed here bit-for-bit (2 for the
      // z taps of that plane + the plane itself)
      const int nl = sl.hl ? 3 : 0, nh = sl.hh ? 3 : 0;
      B2M_TRY(b2m_reserve(ctx, BUF_HALO_LO, 3 * nxy * 4));
      B2M_TRY(b2m_reserve(ctx, BUF_HALO_HI, 3 * nxy * 4));
      float *hlo = b2m_ptr<float>(ctx, BUF_HALO_LO), *hhi = b2m_ptr<float>(ctx, BUF_HALO_HI);
      // NCCL: the exchange runs on a stream of its own while the ctx stream smooths the planes that need own raw
      // planes only; the few planes next to the seams follow once the halo has arrived
      static const bool want_async = !(getenv("B2M_HALO_ASYNC") && atoi(getenv("B2M_HALO_ASYNC")) == 0);
      halo_async = want_async && b2m_comm_async_capable(comm) && sl.nzl >= 16;
      cudaStream_t xs = ctx->stream;
      if (halo_async) {
        B2M_TRY(b2m_aux_stream(ctx));
        xs = ctx->aux_stream;
        CU_TRY(cudaEventRecord(ctx->aux_ev[0], ctx->stream));  // whatever produced d_img / last read the halo buffers
        CU_TRY(cudaStreamWaitEvent(xs, ctx->aux_ev[0], 0));
      }
      B2M_TRY(b2m_comm_exchange_on(ctx, comm, xs, d_img + (size_t)(sl.nzl - 3) * nxy, sl.hh ? 3 * nxy * 4 : 0, hlo, nl * nxy * 4,
                                   d_img, sl.hl ? 3 * nxy * 4 : 0, hhi, nh * nxy * 4));
      if (halo_async) CU_TRY(cudaEventRecord(ctx->aux_ev[1], xs));
      src.lo = hlo; src.n_lo = nl; src.hi = hhi; src.n_hi = nh; src.rz0 = sl.z0 - nl;
    }
    if (!slabs && ctx->pend_n > 0) {
      // the volume is still arriving in z-chunks (b2m_h2d_chunked_begin): smooth the output planes whose raw planes
      // (z-2 .. z+2) are on the device, chunk by chunk, under the rest of the transfer
      int zdone = 0;
      for (int k = 0; k < ctx->pend_n; k++) {
        CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->pend_ev[k], 0));
        const int avail = ctx->pend_zend[k];
        const int zend = avail >= sl.gnz ? sl.gnz : avail - 2;
        if (zend > zdone) {
          src.oz0 = zdone; src.onz = zend - zdone;
          B2M_TRY(smooth_planes(src, (size_t)zdone));
          zdone = zend;
        }
      }
      ctx->pend_n = 0;
    } else if (halo_async) {
      // output planes [ia, ib) read own raw planes only (z-2 .. z+2 inside [z0, z0+nzl), or clipped by the volume)
      const int ia = sl.hl ? sl.z0 + 2 : sl.ez0, ib = sl.hh ? sl.z0 + sl.nzl - 2 : sl.ez0 + sl.nze;
      src.oz0 = ia; src.onz = ib - ia;
      B2M_TRY(smooth_planes(src, (size_t)(ia - sl.ez0)));
      CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[1], 0));
      if (sl.hl) {
        src.oz0 = sl.ez0; src.onz = ia - sl.ez0;
        B2M_TRY(smooth_planes(src, 0));
      }
      if (sl.hh) {
        src.oz0 = ib; src.onz = sl.ez0 + sl.nze - ib;
        B2M_TRY(smooth_planes(src, (size_t)(ib - sl.ez0)));
      }
    } else {
      B2M_TRY(smooth_planes(src, 0));  // range reduction fused (halo planes are other ranks' planes: harmless)
    }
    fo->S = S;
    fo->bits_ready = bits_all;
    fo->bits_iso = sb.iso;
    B2M_TRY(stage_end(ctx, B2M_T_SMOOTH));
  } else {
    if (ctx->pend_n > 0) {  // no smooth to overlap with: wait for the whole volume
      CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->pend_ev[ctx->pend_n - 1], 0));
      ctx->pend_n = 0;
    }
    B2M_TRY(stage_begin(ctx, B2M_T_RANGE));
    if (!slabs) {
      fo->S = d_img;
    } else {
      // EXT copy of the raw slab: one plane from each neighbour around the own planes
      B2M_TRY(b2m_reserve(ctx, BUF_SMOOTH, (size_t)g.n * 4));
      float *S = b2m_ptr<float>(ctx, BUF_SMOOTH);
      CU_TRY(cudaMemcpyAsync(S + (size_t)sl.hl * nxy, d_img, (size_t)sl.nzl * nxy * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      B2M_TRY(b2m_comm_exchange(ctx, comm, d_img + (size_t)(sl.nzl - 1) * nxy, sl.hh ? nxy * 4 : 0, S, sl.hl ? nxy * 4 : 0,
                                d_img, sl.hl ? nxy * 4 : 0, S + (size_t)(sl.hl + sl.nzl) * nxy, sl.hh ? nxy * 4 : 0));
      fo->S = S;
    }
    B2M_TRY(b2m_minmax_run(ctx, d_img, (size_t)sl.nzl * nxy, d_sc));
    B2M_TRY(stage_end(ctx, B2M_T_RANGE));
  }
  B2M_TRY(b2m_sync_scalars(ctx, comm));
  unsigned mn_enc = 0xffffffffu, mx_enc = 0u;
  for (int r = 0; r < sl.world; r++) {
    const b2m_scalars *sr = b2m_sc(ctx, comm, r);
    if (sr->vmin_enc < mn_enc) mn_enc = sr->vmin_enc;
    if (sr->vmax_enc > mx_enc) mx_enc = sr->vmax_enc;
  }
  const float mn = f32_dec(mn_enc), mx = f32_dec(mx_enc);
  const bool talk = sl.rank == 0;  // the reference's stdout lines: once per volume
  if (o->verbose && smooth && talk) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[2 * B2M_T_SMOOTH], ctx->ev[2 * B2M_T_SMOOTH + 1]);
    printf("pre-smooth: %ld ms\n", lroundf(ms));
  }
  if (mn == mx) {  // meshify.c:312-315
    if (talk) printf("Error: No variability in image intensity.\n");
    return B2M_FAIL;
  }
  float iso = o->isolevel;
  fo->iso_reset = 0;
  if (iso <= mn || iso > mx) {  // meshify.c:316-319
    iso = (float)(0.5 * (double)(mn + mx));
    fo->iso_reset = 1;
    if (talk) printf("Suggested isolevel out of range. Intensity range %g..%g, setting isolevel to %g\n", mn, mx, iso);
  }
  if (o->verbose && talk) printf("intensity range %g..%g, isolevel %g\n", mn, mx, iso);
  fo->iso = iso; fo->vmin = mn; fo->vmax = mx;
  fo->edge_max = (float)(0.75 * (double)(mn + iso));  // meshify.c:346: f32 add, double multiply, f32 store

  const bool cc = o->only_largest || o->fill_bubbles;
  B2M_TRY(stage_begin(ctx, cc ? B2M_T_CC : B2M_T_COMPOSE));
  B2M_TRY(b2m_cc_run(ctx, comm, sl, g, o, d_sc, fo));
  B2M_TRY(stage_end(ctx, cc ? B2M_T_CC : B2M_T_COMPOSE));
  B2M_TRY(b2m_sync_scalars(ctx, comm));
  if (ctx->h_scalars->overflow & ~1u) { b2m_set_error("cc: seam list overflow (flags %u)", ctx->h_scalars->overflow); return B2M_ECUDA; }
  if (o->verbose && cc && talk) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[2 * B2M_T_CC], ctx->ev[2 * B2M_T_CC + 1]);
    printf("voxel clustering (largest cluster, bubbles): %ld ms\n", lroundf(ms));
  }
  const int dims[3] = {g.nx, g.ny, sl.gnz};
  for (int a = 0; a < 3; a++) {  // meshify.c:368-371
    int lo = INT_MAX, hi = -1;
    for (int r = 0; r < sl.world; r++) {
      const b2m_scalars *sr = b2m_sc(ctx, comm, r);
      if (sr->lo[a] < lo) lo = sr->lo[a];
      if (sr->hi[a] > hi) hi = sr->hi[a];
    }
    if (hi < 0) { lo = dims[a]; hi = 0; }  // no bright voxel at all (cannot happen: iso <= max)
    fo->lo[a] = lo - 1 > 0 ? lo - 1 : 0;
    fo->hi[a] = hi + 2 < dims[a] ? hi + 2 : dims[a];
  }
  if (res) {
    res->iso_used = iso; res->vmin = mn; res->vmax = mx; res->iso_reset = fo->iso_reset;
    for (int a = 0; a < 3; a++) { res->lo[a] = fo->lo[a]; res->hi[a] = fo->hi[a]; }
  }
  return B2M_OK;
}

// the whole path on one slab (world == 1: the whole volume).  wo = this rank's part of the welded mesh.
static int meshify_slab_impl(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const float *d_img, const int64_t gdims[3],
                             const b2m_opts *o, b2m_result *res, b2m_weld_out *wo) {
  const int64_t edims[3] = {gdims[0], gdims[1], sl.nze};
  b2m_geom g = b2m_make_geom(edims);
  b2m_front_out fo;
  memset(&fo, 0, sizeof(fo));
  memset(wo, 0, sizeof(*wo));
  const bool talk = sl.rank == 0;
  B2M_TRY(stage_begin(ctx, B2M_T_TOTAL));
  B2M_TRY(b2m_front_run(ctx, comm, sl, d_img, g, o, &fo, res));
  b2m_mesh_dev mesh;
  B2M_TRY(stage_begin(ctx, B2M_T_MC));
  int rc = b2m_mc_run(ctx, comm, sl, g, o, &fo, &mesh);
  if (rc != B2M_OK) {
    if (rc == B2M_FAIL && o->backend == B2M_BACKEND_CLASSIC && talk)
      printf("marching cubes failed to identify triangles with an isolevel of %g\n", fo.iso);
    return rc;
  }
  B2M_TRY(stage_end(ctx, B2M_T_MC));
  res->pre_nverts = o->backend == B2M_BACKEND_CLASSIC ? (int)(3 * mesh.NT) : (int)(mesh.NVE + mesh.NVC);
  res->pre_ntris = (int)mesh.NT;
  if (o->verbose && talk) {
    float ms = 0.f;
    cudaEventSynchronize(ctx->ev[2 * B2M_T_MC + 1]);
    cudaEventElapsedTime(&ms, ctx->ev[2 * B2M_T_MC], ctx->ev[2 * B2M_T_MC + 1]);
    printf("marching cubes (%dx%dx%d): %ld ms\n", g.nx, g.ny, sl.gnz, lroundf(ms));
  }
  B2M_TRY(b2m_weld_run(ctx, comm, &mesh, 0, wo));
  ctx->ev_mask |= (1u << B2M_T_WELD) | (1u << B2M_T_DEGEN);
  B2M_TRY(stage_end(ctx, B2M_T_TOTAL));
  // the last host round trip of the call: weld bases, surviving triangles (of every rank), consistency flags
  B2M_TRY(b2m_sync_scalars(ctx, comm));
  B2M_TRY(b2m_weld_finish(ctx, comm, &mesh, wo));
  unsigned long long NTk = 0;
  unsigned t_off_new = 0;
  for (int r = 0; r < sl.world; r++) {
    if (r == sl.rank) t_off_new = (unsigned)NTk;
    NTk += b2m_sc(ctx, comm, r)->n_tri_kept;
  }
  B2M_TRY(collect_times(ctx, res));
  res->nverts = (int)wo->nv_global;
  res->ntris = (int)NTk;
  res->nmerged = (int)(wo->n_dead) - (int)(wo->n_extra);
  res->ndegenerate = (int)(mesh.NT - NTk);
  res->d_verts = wo->verts;
  res->d_tris = wo->tris;
  ctx->slab_t_off = t_off_new;
  if (o->verbose && talk) {
    if (wo->n_dead) printf("vertex welding %d -> %d: %ld ms\n", res->pre_nverts, res->nverts, lroundf(res->ms[B2M_T_WELD]));
    else printf("Unify vertices found no shared vertices\n");
    if (res->ndegenerate)
      printf("remove degenerate triangles %d -> %d: %ld ms\n", res->pre_ntris, res->ntris, lroundf(res->ms[B2M_T_DEGEN]));
  }
  if (res->nverts < 3) return B2M_FAIL;  // meshify.c:382
  return B2M_OK;
}

extern "C" int b2m_meshify_device(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const b2m_opts *opts,
                                  b2m_result *res) {
  if (!ctx || !d_img || !dims || !opts || !res) { b2m_set_error("null argument"); return B2M_EARG; }
  B2M_TRY(check_dims(dims));
  CU_TRY(cudaSetDevice(ctx->device));
  memset(res, 0, sizeof(*res));
  ctx->launches = 0;
  ctx->ev_mask = 0;
  ctx->nkt = 0;
  b2m_weld_out wo;
  int rc = meshify_slab_impl(ctx, nullptr, single_slab((int)dims[2]), d_img, dims, opts, res, &wo);
  if (rc != B2M_OK) cudaStreamSynchronize(ctx->stream);
  return rc;
}

// z-slab entry point: every rank of `comm` calls it with its own planes [z0, z0+nzl) of a volume of gdims.
extern "C" int b2m_meshify_slab(b2m_ctx *ctx, b2m_comm *comm, const float *d_slab, const int64_t gdims[3], int64_t z0,
                                int64_t nzl, const b2m_opts *opts, b2m_slab_result *out) {
  if (!ctx || !d_slab || !gdims || !opts || !out) { b2m_set_error("null argument"); return B2M_EARG; }
  const int W = b2m_comm_world(comm), rank = b2m_comm_rank(comm);
  int rc = B2M_OK;
  for (int a = 0; a < 3; a++)
    if (gdims[a] < 1 || gdims[a] > 32767) { b2m_set_error("gdims[%d] outside 1..32767", a); rc = B2M_EARG; }
  if (rc == B2M_OK && (z0 < 0 || nzl < 1 || z0 + nzl > gdims[2] || (rank == 0) != (z0 == 0) || (rank == W - 1) != (z0 + nzl == gdims[2]))) {
    b2m_set_error("slab [%lld, %lld) of rank %d/%d does not fit a volume of %lld planes", (long long)z0, (long long)(z0 + nzl), rank, W, (long long)gdims[2]);
    rc = B2M_EARG;
  }
  if (rc == B2M_OK && W > 1 && nzl < 4) { b2m_set_error("slabs need at least 4 planes each"); rc = B2M_EARG; }
  // a slab holds at most 2^32 voxels (2^27 bit words: run slots stay below 2^31), the whole volume whatever the ranks
  // can hold; mesh counts are checked against the int counts of the API where they are known (b2m_mc_run)
  if (rc == B2M_OK && ((unsigned long long)((gdims[0] + 31) / 32) * gdims[1] * nzl > (1ull << 27))) {
    b2m_set_error("slab of more than 2^27 bit words (2^32 voxels)");
    rc = B2M_EARG;
  }
  if (rc != B2M_OK) { b2m_comm_abort(comm); return rc; }
  CU_TRY(cudaSetDevice(ctx->device));
  memset(out, 0, sizeof(*out));
  ctx->launches = 0;
  ctx->ev_mask = 0;
  ctx->nkt = 0;
  b2m_slab sl;
  sl.rank = rank; sl.world = W; sl.gnz = (int)gdims[2]; sl.z0 = (int)z0; sl.nzl = (int)nzl;
  sl.hl = rank > 0; sl.hh = rank < W - 1; sl.ez0 = sl.z0 - sl.hl; sl.nze = sl.nzl + sl.hl + sl.hh;
  b2m_weld_out wo;
  rc = meshify_slab_impl(ctx, comm, sl, d_slab, gdims, opts, &out->r, &wo);
  if (rc != B2M_OK) {
    cudaStreamSynchronize(ctx->stream);
    if (rc != B2M_FAIL) b2m_comm_abort(comm);  // B2M_FAIL is decided identically on every rank
    return rc;
  }
  out->d_verts = wo.verts; out->d_tris = wo.tris;
  out->nv_edge = (int)wo.nve_local; out->nv_cent = (int)wo.nvc_local; out->nv_extra = (int)wo.nx_local;
  out->ntris_local = (int)wo.nt_local;
  out->v_edge_off = wo.v_edge_off; out->v_cent_off = wo.v_c_off;
  out->v_extra_off = (int64_t)wo.nv_global - wo.n_extra;
  out->tri_off = ctx->slab_t_off;
  return B2M_OK;
}

extern "C" int b2m_fetch_mesh(b2m_ctx *ctx, const b2m_result *res, void *h_verts, void *h_tris) {
  if (!ctx || !res) return B2M_EARG;
  CU_TRY(cudaSetDevice(ctx->device));
  if (h_verts && res->nverts) B2M_TRY(b2m_copy_d2h(ctx, h_verts, res->d_verts, (size_t)res->nverts * 24));
  if (h_tris && res->ntris) B2M_TRY(b2m_copy_d2h(ctx, h_tris, res->d_tris, (size_t)res->ntris * 12));
  return B2M_OK;
}

// output blocks allocated as soon as the marching-cubes totals are known, pre-faulted on the copy pool while the
// emit / weld kernels run (first-touch faults of fresh malloc() pages were half of the D2H time)
struct host_out {
  void *v, *t;
  size_t cap_v, cap_t;  // bytes
  size_t seen_v, seen_t;  // marching-cubes totals of this call (remembered for the next call on the same geometry)
  size_t pcie_bytes;      // bytes the D2H copies moved
  int touching;
};
static void host_out_hook(void *user, size_t nverts, size_t ntris);
// A repeated call on a volume of the same size starts allocating / pre-faulting its output blocks right away, from
// the previous call's totals, so the page faults overlap the H2D copy as well (wrong guesses are re-allocated).
static void host_out_early(b2m_ctx *ctx, host_out *h, size_t nvox) {
  if (ctx->last_nvox == nvox && ctx->last_nv) host_out_hook(h, ctx->last_nv + ctx->last_nv / 64, ctx->last_nt + ctx->last_nt / 64);
  h->seen_v = h->seen_t = 0;
}
static void host_out_remember(b2m_ctx *ctx, const host_out *h, size_t nvox) {
  ctx->last_nvox = nvox; ctx->last_nv = h->seen_v; ctx->last_nt = h->seen_t;
}
static void host_out_hook(void *user, size_t nverts, size_t ntris) {
  host_out *h = (host_out *)user;
  h->seen_v = nverts; h->seen_t = ntris;
  if (h->v || h->t) return;
  h->cap_v = (nverts + 4096) * 24 + 8;  // a little head-room for the split-off vertices of the classic weld
  h->cap_t = ntris * 12 + 8;
  h->v = malloc(h->cap_v);
  h->t = malloc(h->cap_t);
  if (!h->v || !h->t) return;
  hint_hugepages(h->v, h->cap_v);  // fresh mmap'd blocks: 2 MB pages cut the first-touch faults 512x
  hint_hugepages(h->t, h->cap_t);
  h->touching = b2m_touch_async(h->v, nverts * 24, h->t, ntris * 12);
}
static int host_out_finish(b2m_ctx *ctx, host_out *h, int rc, const void *d_verts, size_t nv, const void *d_tris, size_t nt,
                           void **verts, void **tris) {
  ctx->counts_hook = nullptr;
  if (h->touching) b2m_touch_wait();
  if (rc == B2M_OK) {
    if (!h->v || h->cap_v < nv * 24 + 8) { free(h->v); h->v = malloc(nv * 24 + 8); if (h->v) hint_hugepages(h->v, nv * 24); }
    if (!h->t || h->cap_t < nt * 12 + 8) { free(h->t); h->t = malloc(nt * 12 + 8); if (h->t) hint_hugepages(h->t, nt * 12); }
    if (!h->v || !h->t) { b2m_set_error("malloc of the output mesh failed"); rc = B2M_ENOMEM; }
  }
  if (rc == B2M_OK && nv) {
    // Lewiner positions are f32 values widened to f64: 12 instead of 24 bytes per vertex cross PCIe (checked on the
    // device; anything else - the classic back-end's FP64 positions - takes the plain copy)
    // (in few-threads mode the doubles go straight into the registered block: one pass over host DRAM beats half the
    // PCIe bytes when the node's memory bandwidth is the bound)
    int done = 0;
    if (!(b2m_d2h_registers() && nv * 24 >= ((size_t)64 << 20)))
      rc = b2m_copy_d2h_f32exact(ctx, (double *)h->v, (const double *)d_verts, nv * 3, &done);
    if (rc == B2M_OK && !done) rc = b2m_copy_d2h(ctx, h->v, d_verts, nv * 24);
    h->pcie_bytes = nv * (done ? 12 : 24);
  }
  if (rc == B2M_OK && nt) rc = b2m_copy_d2h(ctx, h->t, d_tris, nt * 12);
  h->pcie_bytes += nt * 12;
  if (rc != B2M_OK) { free(h->v); free(h->t); return rc; }
  *verts = h->v;
  *tris = h->t;
  return B2M_OK;
}

extern "C" int b2m_meshify_host(b2m_ctx *ctx, const float *h_img, const int64_t dims[3], const b2m_opts *opts,
                                void **verts, void **tris, b2m_result *res) {
  if (!ctx || !h_img || !dims || !opts || !res || !verts || !tris) { b2m_set_error("null argument"); return B2M_EARG; }
  B2M_TRY(check_dims(dims));
  CU_TRY(cudaSetDevice(ctx->device));
  size_t n = (size_t)dims[0] * dims[1] * dims[2];
  B2M_TRY(b2m_reserve(ctx, BUF_INPUT, n * 4));
  host_out ho;
  memset(&ho, 0, sizeof(ho));
  host_out_early(ctx, &ho, n);
  const double t0 = wall_ms();
  // a pinned volume goes up in z-chunks on a second stream and the smooth follows it (B2M_H2D_OVERLAP=0: one plain copy)
  static const bool want_overlap = !(getenv("B2M_H2D_OVERLAP") && atoi(getenv("B2M_H2D_OVERLAP")) == 0);
  const bool overlap = want_overlap && opts->pre_smooth && dims[0] >= 5 && dims[1] >= 5 && dims[2] >= 64 && n * 4 >= ((size_t)256 << 20) &&
                       b2m_host_is_pinned(h_img);
  int rc = overlap ? b2m_h2d_chunked_begin(ctx, b2m_ptr<float>(ctx, BUF_INPUT), h_img, (size_t)dims[0] * dims[1], (int)dims[2])
                   : b2m_copy_h2d(ctx, ctx->buf[BUF_INPUT].p, h_img, n * 4);
  const double t1 = wall_ms();
  ctx->counts_hook = host_out_hook;
  ctx->hook_user = &ho;
  if (rc == B2M_OK) rc = b2m_meshify_device(ctx, b2m_ptr<float>(ctx, BUF_INPUT), dims, opts, res);
  if (overlap) {  // whatever happened, the transfer must have drained before the buffers are reused
    ctx->pend_n = 0;
    cudaStreamSynchronize(ctx->copy_stream);
  }
  const double t2 = wall_ms();
  host_out_remember(ctx, &ho, n);
  rc = host_out_finish(ctx, &ho, rc, res->d_verts, (size_t)res->nverts, res->d_tris, (size_t)res->ntris, verts, tris);
  if (rc != B2M_OK) return rc;
  res->h2d_ms = (float)(t1 - t0);
  if (overlap) b2m_h2d_chunked_ms(ctx, &res->h2d_ms);  // the transfer's own duration; ms[] (device) includes waiting for it
  res->d2h_ms = (float)(wall_ms() - t2);
  res->d2h_bytes = ho.pcie_bytes;
  if (opts->verbose) printf("host copies: H2D %.1f ms, D2H %.1f ms\n", res->h2d_ms, res->d2h_ms);
  return B2M_OK;
}

// meshify() from the RAW voxels of a NIfTI file and with the voxel->world transform fused in: what nii2() does around
// meshify() (load_nii's type conversion, src/nii2mesh.c:155-172; apply_sform, src/nii2mesh.c:328), minus the 4-byte
// widening before the PCIe copy and the host pass over the vertices.  srow_* may be NULL (voxel coordinates).
extern "C" int b2m_ingest_host(b2m_ctx *ctx, const void *h_raw, int datatype, size_t nvox, float scl_slope, float scl_inter, float *d_out);
extern "C" int b2m_apply_sform_device(b2m_ctx *ctx, const b2m_result *res, const float srow_x[4], const float srow_y[4], const float srow_z[4]);
extern "C" int b2m_meshify_raw_host(b2m_ctx *ctx, const void *h_raw, int datatype, const int64_t dims[3], float scl_slope,
                                    float scl_inter, const b2m_opts *opts, const float *srow_x, const float *srow_y,
                                    const float *srow_z, void **verts, void **tris, b2m_result *res) {
  if (!ctx || !h_raw || !dims || !opts || !res || !verts || !tris) { b2m_set_error("null argument"); return B2M_EARG; }
  B2M_TRY(check_dims(dims));
  CU_TRY(cudaSetDevice(ctx->device));
  const size_t n = (size_t)dims[0] * dims[1] * dims[2];
  B2M_TRY(b2m_reserve(ctx, BUF_INPUT, n * 4));
  host_out ho;
  memset(&ho, 0, sizeof(ho));
  host_out_early(ctx, &ho, n);
  const double t0 = wall_ms();
  int rc = b2m_ingest_host(ctx, h_raw, datatype, n, scl_slope, scl_inter, b2m_ptr<float>(ctx, BUF_INPUT));
  const double t1 = wall_ms();
  ctx->counts_hook = host_out_hook;
  ctx->hook_user = &ho;
  if (rc == B2M_OK) rc = b2m_meshify_device(ctx, b2m_ptr<float>(ctx, BUF_INPUT), dims, opts, res);
  if (rc == B2M_OK && srow_x && srow_y && srow_z) rc = b2m_apply_sform_device(ctx, res, srow_x, srow_y, srow_z);
  const double t2 = wall_ms();
  host_out_remember(ctx, &ho, n);
  rc = host_out_finish(ctx, &ho, rc, res->d_verts, (size_t)res->nverts, res->d_tris, (size_t)res->ntris, verts, tris);
  if (rc != B2M_OK) return rc;
  res->h2d_ms = (float)(t1 - t0);
  res->d2h_ms = (float)(wall_ms() - t2);
  res->d2h_bytes = ho.pcie_bytes;
  return B2M_OK;
}

// z-slab path from HOST memory: this rank's planes in, this rank's mesh blocks out (malloc()'d, caller frees).
extern "C" int b2m_meshify_slab_host(b2m_ctx *ctx, b2m_comm *comm, const float *h_slab, const int64_t gdims[3], int64_t z0,
                                     int64_t nzl, const b2m_opts *opts, void **verts, void **tris, b2m_slab_result *out) {
  if (!ctx || !h_slab || !gdims || !opts || !out || !verts || !tris) { b2m_set_error("null argument"); b2m_comm_abort(comm); return B2M_EARG; }
  if (nzl < 1 || gdims[0] < 1 || gdims[1] < 1) { b2m_set_error("bad slab"); b2m_comm_abort(comm); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  const size_t n = (size_t)gdims[0] * gdims[1] * nzl;
  int rc = b2m_reserve(ctx, BUF_INPUT, n * 4);
  host_out ho;
  memset(&ho, 0, sizeof(ho));
  if (rc == B2M_OK) host_out_early(ctx, &ho, n);
  const double t0 = wall_ms();
  if (rc == B2M_OK) rc = b2m_copy_h2d(ctx, ctx->buf[BUF_INPUT].p, h_slab, n * 4);
  if (rc != B2M_OK) {
    b2m_comm_abort(comm);
    if (ho.touching) b2m_touch_wait();
    free(ho.v); free(ho.t);
    return rc;
  }
  const double t1 = wall_ms();
  ctx->counts_hook = host_out_hook;
  ctx->hook_user = &ho;
  rc = b2m_meshify_slab(ctx, comm, b2m_ptr<float>(ctx, BUF_INPUT), gdims, z0, nzl, opts, out);
  const double t2 = wall_ms();
  host_out_remember(ctx, &ho, n);
  const size_t nv = (size_t)out->nv_edge + out->nv_cent + out->nv_extra, nt = (size_t)out->ntris_local;
  rc = host_out_finish(ctx, &ho, rc, out->d_verts, nv, out->d_tris, nt, verts, tris);
  if (rc != B2M_OK) return rc;
  out->r.h2d_ms = (float)(t1 - t0);
  out->r.d2h_ms = (float)(wall_ms() - t2);
  out->r.d2h_bytes = ho.pcie_bytes;
  return B2M_OK;
}

// ---- stage hooks ---------------------------------------------------------------------------------
extern "C" int b2m_stage_smooth(b2m_ctx *ctx, const float *d_in, float *d_out, const int64_t dims[3]) {
  if (!ctx || !d_in || !d_out || !dims) return B2M_EARG;
  B2M_TRY(check_dims(dims));
  CU_TRY(cudaSetDevice(ctx->device));
  b2m_geom g = b2m_make_geom(dims);
  if (g.nx < 5 || g.ny < 5 || g.nz < 5) {
    CU_TRY(cudaMemcpyAsync(d_out, d_in, (size_t)g.n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return B2M_FAIL;
  }
  B2M_TRY(init_scalars(ctx));
  b2m_scalars *d_sc = b2m_ptr<b2m_scalars>(ctx, BUF_SCALARS);
  smooth_src src;
  memset(&src, 0, sizeof(src));
  src.lo = src.main = src.hi = d_in;
  src.n_main = g.nz; src.gnz = g.nz; src.onz = g.nz;
  B2M_TRY(b2m_smooth_run(ctx, src, d_out, g, d_sc));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B2M_OK;
}

extern "C" int b2m_stage_front(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const b2m_opts *opts,
                               float *d_composed, uint8_t *d_mask, b2m_result *res) {
  if (!ctx || !d_img || !dims || !opts || !res) return B2M_EARG;
  B2M_TRY(check_dims(dims));
  CU_TRY(cudaSetDevice(ctx->device));
  memset(res, 0, sizeof(*res));
  ctx->launches = 0;
  ctx->ev_mask = 0;
  ctx->nkt = 0;
  b2m_geom g = b2m_make_geom(dims);
  b2m_front_out fo;
  memset(&fo, 0, sizeof(fo));
  B2M_TRY(b2m_front_run(ctx, nullptr, single_slab(g.nz), d_img, g, opts, &fo, res));
  if (d_composed || d_mask)
    B2M_TRY(b2m_compose_materialize(ctx, g, &fo, d_composed, d_mask, b2m_ptr<b2m_scalars>(ctx, BUF_SCALARS), 0));
  B2M_TRY(collect_times(ctx, res));
  return B2M_OK;
}

extern "C" int b2m_stage_mc(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const int lo[3], const int hi[3],
                            const b2m_opts *opts, b2m_result *res) {
  if (!ctx || !d_img || !dims || !opts || !res || !lo || !hi) return B2M_EARG;
  B2M_TRY(check_dims(dims));
  CU_TRY(cudaSetDevice(ctx->device));
  memset(res, 0, sizeof(*res));
  ctx->launches = 0;
  ctx->ev_mask = 0;
  ctx->nkt = 0;
  b2m_geom g = b2m_make_geom(dims);
  B2M_TRY(init_scalars(ctx));
  b2m_scalars *d_sc = b2m_ptr<b2m_scalars>(ctx, BUF_SCALARS);
  b2m_scalars *h = ctx->h_scalars;
  // the volume is taken as already composed: no fill/keep masks, no darkening (edge_max = +inf);
  // the pad value needs the true minimum (src/MarchingCubes.c:1097-1100)
  B2M_TRY(b2m_minmax_run(ctx, d_img, (size_t)g.n, d_sc));
  B2M_TRY(b2m_fetch_scalars(ctx));
  b2m_front_out fo;
  memset(&fo, 0, sizeof(fo));
  fo.S = d_img;
  fo.iso = opts->isolevel;
  fo.vmin = f32_dec(h->vmin_enc);
  fo.vmax = f32_dec(h->vmax_enc);
  fo.edge_max = INFINITY;
  for (int a = 0; a < 3; a++) {
    if (lo[a] < 0 || hi[a] > dims[a] || lo[a] > hi[a]) { b2m_set_error("bad bbox"); return B2M_EARG; }
    fo.lo[a] = lo[a]; fo.hi[a] = hi[a];
  }
  // inside bits of the (already composed) volume: the marching-cubes comparison alone
  B2M_TRY(b2m_reserve(ctx, BUF_FG, (size_t)g.nwords * 4));
  B2M_TRY(b2m_reserve(ctx, BUF_MB, (size_t)g.nwords * 4));
  B2M_TRY(b2m_threshold_run(ctx, d_img, g, fo.iso, b2m_ptr<uint32_t>(ctx, BUF_FG), nullptr, b2m_ptr<uint32_t>(ctx, BUF_MB),
                            opts->backend == B2M_BACKEND_CLASSIC));
  fo.ibits = b2m_ptr<uint32_t>(ctx, BUF_MB);
  b2m_mesh_dev mesh;
  memset(&mesh, 0, sizeof(mesh));
  B2M_TRY(stage_begin(ctx, B2M_T_MC));
  int rc = b2m_mc_run(ctx, nullptr, single_slab(g.nz), g, opts, &fo, &mesh);
  if (rc != B2M_OK) { cudaStreamSynchronize(ctx->stream); return rc; }
  B2M_TRY(stage_end(ctx, B2M_T_MC));
  res->nverts = (int)(mesh.nv_edge + mesh.nv_c); res->ntris = (int)mesh.nt;
  res->pre_nverts = res->nverts; res->pre_ntris = (int)mesh.nt;
  res->d_verts = mesh.verts; res->d_tris = mesh.tris;
  res->iso_used = fo.iso; res->vmin = fo.vmin; res->vmax = fo.vmax;
  B2M_TRY(collect_times(ctx, res));
  return B2M_OK;
}

extern "C" int b2m_stage_weld(b2m_ctx *ctx, double *h_verts, int *h_tris, int *nv, int *nt) {
  if (!ctx || !h_verts || !h_tris || !nv || !nt || *nv < 1 || *nt < 1) return B2M_EARG;
  CU_TRY(cudaSetDevice(ctx->device));
  ctx->launches = 0;
  ctx->ev_mask = 0;
  ctx->nkt = 0;
  B2M_TRY(init_scalars(ctx));
  B2M_TRY(b2m_reserve(ctx, BUF_VERTS, (size_t)*nv * 24));
  B2M_TRY(b2m_reserve(ctx, BUF_TRIS, (size_t)*nt * 12));
  CU_TRY(cudaMemcpyAsync(ctx->buf[BUF_VERTS].p, h_verts, (size_t)*nv * 24, cudaMemcpyHostToDevice, ctx->stream));
  CU_TRY(cudaMemcpyAsync(ctx->buf[BUF_TRIS].p, h_tris, (size_t)*nt * 12, cudaMemcpyHostToDevice, ctx->stream));
  b2m_mesh_dev mesh;
  memset(&mesh, 0, sizeof(mesh));
  mesh.verts = b2m_ptr<double>(ctx, BUF_VERTS);
  mesh.tris = b2m_ptr<int>(ctx, BUF_TRIS);
  mesh.nv_edge = mesh.NVE = (unsigned)*nv;
  mesh.nt = mesh.NT = (unsigned)*nt;
  mesh.d_p0 = mesh.verts;  // the reference's pts[0]
  b2m_weld_out wo;
  memset(&wo, 0, sizeof(wo));
  B2M_TRY(b2m_weld_run(ctx, nullptr, &mesh, 1, &wo));
  B2M_TRY(b2m_fetch_scalars(ctx));
  B2M_TRY(b2m_weld_finish(ctx, nullptr, &mesh, &wo));
  b2m_result res;
  memset(&res, 0, sizeof(res));
  res.d_verts = wo.verts; res.d_tris = wo.tris; res.nverts = (int)wo.nv_local; res.ntris = (int)wo.nt_local;
  B2M_TRY(b2m_fetch_mesh(ctx, &res, h_verts, h_tris));
  *nv = res.nverts;
  *nt = res.ntris;
  return B2M_OK;
}
