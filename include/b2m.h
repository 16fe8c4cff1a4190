/* b2m.h — thin C ABI of the B200 voxel->mesh engine (libb2m.so).
 *
 * Plain pointers and sizes only; no CUDA or torch types in any signature.  Every entry point
 * returns 0 on success, B2M_FAIL (1, the reference's EXIT_FAILURE) when the reference would have
 * failed for the same input, or a negative B2M_E* code for CUDA/usage errors (b2m_last_error()
 * gives the text).  There is NO CPU fallback: without a working CUDA device every call fails.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference/):
 *   b2m_meshify_device / b2m_meshify_host   meshify()                     src/meshify.c:286-389
 *   b2m_stage_smooth                        quick_smooth()                src/meshify.c:170-216
 *   b2m_stage_front                         smooth..edge-darken+bbox      src/meshify.c:299-371
 *                                           (bwlabel() src/bwlabel.c:478-543, dilate() src/meshify.c:218-264)
 *   b2m_stage_mc                            marchingCubes()               src/MarchingCubes.c:1086-1143
 *                                                                         src/oldcubes.c:465-522
 *   b2m_stage_weld                          unify_vertices() +            src/meshify.c:45-106
 *                                           remove_degenerate_triangles() src/meshify.c:113-168
 */
#ifndef B2M_H
#define B2M_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2M_OK 0
#define B2M_FAIL 1          /* reference EXIT_FAILURE semantics (no variability, empty mesh, ...) */
#define B2M_ECUDA (-1)      /* CUDA runtime error */
#define B2M_EARG (-2)       /* bad argument (dims out of range, NULL pointers, ...) */
#define B2M_ENOMEM (-3)     /* device or host allocation failed */

/* marching-cubes back-end: the reference picks this at COMPILE time (src/meshify.c:25-29) */
#define B2M_BACKEND_LEWINER 0 /* MarchingCubes.c: MC33, or its classic table when original_mc != 0 */
#define B2M_BACKEND_CLASSIC 1 /* oldcubes.c: FP64 triangle soup semantics; original_mc ignored */

typedef struct b2m_ctx b2m_ctx; /* one per host thread / stream: device, stream, workspace arena */

typedef struct {
  float isolevel;   /* meshify(isolevel)                                                   */
  int original_mc;  /* meshify(originalMC): -o                                             */
  int pre_smooth;   /* -p                                                                  */
  int only_largest; /* -l                                                                  */
  int fill_bubbles; /* -b                                                                  */
  int backend;      /* B2M_BACKEND_*                                                       */
  int verbose;      /* print the reference's stage lines (with device times) to stdout     */
} b2m_opts;

#define B2M_NSTAGE 8
enum { B2M_T_SMOOTH = 0, B2M_T_RANGE, B2M_T_CC, B2M_T_COMPOSE, B2M_T_MC, B2M_T_WELD, B2M_T_DEGEN, B2M_T_TOTAL };

typedef struct {
  /* mesh, left resident on the device (owned by the ctx; valid until the next call on it) */
  const void *d_verts; /* nverts x 3 f64 (vec3d layout) */
  const void *d_tris;  /* ntris x 3 i32 (vec3i layout)  */
  int nverts, ntris;         /* after weld / degenerate-triangle removal  */
  int pre_nverts, pre_ntris; /* marching-cubes output (Lewiner: as the reference; classic: soup = 3*pre_ntris) */
  int nmerged, ndegenerate;  /* vertices merged by the weld, triangles removed */
  float iso_used;            /* isolevel after the range sanity reset (src/meshify.c:316-319) */
  float vmin, vmax;          /* intensity range after smoothing (src/meshify.c:306-311) */
  int lo[3], hi[3];          /* bright bounding box handed to marching cubes (src/meshify.c:368-371) */
  int iso_reset;             /* 1 if the isolevel was out of range and reset */
  float ms[B2M_NSTAGE];      /* CUDA-event device time per stage, ms; ms[B2M_T_TOTAL] = whole call */
  uint64_t launches;         /* kernels launched by this call */
  float h2d_ms, d2h_ms;      /* b2m_meshify_host only: wall-clock time of the host->device / device->host copies */
  unsigned long long d2h_bytes; /* ... and the bytes the device->host copies moved over PCIe (Lewiner vertices travel as f32) */
} b2m_result;

/* ---- context ------------------------------------------------------------------------------ */
int b2m_create(b2m_ctx **ctx, int device);
void b2m_destroy(b2m_ctx *ctx);
const char *b2m_last_error(void);
const char *b2m_version(void);
void b2m_set_default_backend(int backend); /* used by meshify(); overrides B2M_CLASSIC_CUBES */
int b2m_device_count(void);

/* ---- device memory helpers (so that C or ctypes callers need no CUDA bindings) ------------- */
int b2m_dev_alloc(void **dptr, size_t bytes);  /* on the calling thread's CURRENT device */
int b2m_ctx_alloc(b2m_ctx *ctx, void **dptr, size_t bytes);  /* on the context's device (processes that drive several GPUs) */
int b2m_dev_free(void *dptr);
int b2m_host_alloc(void **hptr, size_t bytes); /* pinned */
int b2m_host_free(void *hptr);
int b2m_h2d(b2m_ctx *ctx, void *dst, const void *src, size_t bytes);
int b2m_d2h(b2m_ctx *ctx, void *dst, const void *src, size_t bytes);
int b2m_sync(b2m_ctx *ctx);
int b2m_flush_l2(b2m_ctx *ctx); /* writes a 256 MiB scratch buffer (benchmark hygiene) */
/* Host threads of the copy pool that moves pageable host memory to / from the pinned staging buffers (the caller
 * included).  Default: B2M_COPY_THREADS, else (cores the process may run on) / LOCAL_WORLD_SIZE (as exported by torchrun:
 * the ranks that share this node share its cores), at most 16.  b2m_set_copy_threads(n) must precede the first bulk copy
 * of the process (returns B2M_FAIL afterwards); n = 0 restores the default rule. */
int b2m_set_copy_threads(int n);
int b2m_get_copy_threads(void);
int b2m_pool_selftest(size_t bytes, int callers); /* the pool alone, no GPU: concurrent callers with a barrier between them */
/* d_out[z][y][x] = d_tile[(z+z_offset) % tz][y % ty][x % tx]: periodic replication of a small device
 * tile into a (slab of a) large volume without a host copy (synthetic G1024 / G2048 inputs). */
int b2m_tile_volume(b2m_ctx *ctx, const float *d_tile, const int64_t tile_dims[3], float *d_out,
                    const int64_t dims[3], int64_t z_offset);

/* ---- measurement --------------------------------------------------------------------------- */
/* per-kernel device times of the LAST hot-path call on this ctx: when profiling is on, every kernel
 * launch is bracketed by a CUDA-event pair on the ctx stream (the live launch list of bench.py). */
int b2m_set_profile(b2m_ctx *ctx, int on);
int b2m_profile_count(b2m_ctx *ctx);
int b2m_profile_entry(b2m_ctx *ctx, int i, const char **name, float *ms);
/* one CUDA-event pair on the ctx stream around an arbitrary region of calls */
int b2m_timer_start(b2m_ctx *ctx);
int b2m_timer_stop(b2m_ctx *ctx, float *ms);

/* ---- the hot path -------------------------------------------------------------------------- */
/* Whole meshify() on a volume already resident in device memory (x fastest, dims = NX,NY,NZ).
 * d_img is NOT modified.  Results stay on the device (res->d_verts / d_tris).  This is what the
 * Gvoxel/s metric times (res->ms[B2M_T_TOTAL], CUDA events on the ctx stream). */
int b2m_meshify_device(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const b2m_opts *opts,
                       b2m_result *res);

/* Same, from a HOST volume: H2D copy, device pipeline, D2H into malloc() blocks (*verts: nverts x
 * 3 f64, *tris: ntris x 3 i32) that the caller free()s.  This is what meshify() calls. */
int b2m_meshify_host(b2m_ctx *ctx, const float *h_img, const int64_t dims[3], const b2m_opts *opts,
                     void **verts, void **tris, b2m_result *res);

/* ---- z-slabs across GPUs (replaces nothing in the reference, which is single-threaded) ----
 * The volume is cut into contiguous z-slabs, one rank (GPU) per slab, raster order = rank order.  Every rank
 * calls b2m_meshify_slab() with its own planes; the ranks exchange halo planes for the smooth and for marching
 * cubes, merge connected components across the slab faces, number the vertices globally in the reference's
 * emission order and weld the seam vertices.  The assembled mesh equals b2m_meshify_device() on the whole
 * volume bit for bit.  Transport: NCCL (one process per GPU, libnccl.so.2 loaded at run time) or a
 * single-process group of host threads (peer copies). */
typedef struct b2m_comm b2m_comm;
int b2m_comm_nccl_id(void *id128);  /* rank 0: 128-byte ncclUniqueId to hand to the other ranks */
int b2m_comm_create_nccl(b2m_comm **comm, b2m_ctx *ctx, const void *id128, int rank, int world);
int b2m_comm_create_local(b2m_comm **comms /* [world] */, int world);  /* one handle per host thread */
void b2m_comm_destroy(b2m_comm *comm);
int b2m_comm_reset(b2m_comm *comm);

typedef struct {
  b2m_result r;          /* GLOBAL counts (nverts, ntris, pre_*), range, isolevel, bbox, this rank's stage times;
                            r.d_verts / r.d_tris = this rank's blocks */
  const void *d_verts;   /* this rank's welded vertices: [edge block | centroid block | extra block], 3 f64 each */
  const void *d_tris;    /* this rank's triangles, 3 i32 each, indices into the ASSEMBLED vertex array */
  int nv_edge, nv_cent, nv_extra, ntris_local;
  int64_t v_edge_off, v_cent_off, v_extra_off;  /* where the three blocks sit in the assembled vertex array */
  int64_t tri_off;                              /* where the triangles sit in the assembled triangle array */
} b2m_slab_result;

/* d_slab: planes [z0, z0+nzl) of a volume of gdims (x fastest), resident on ctx's device; comm may be NULL for a
 * single slab covering the whole volume.  Collective: every rank of comm must call it. */
int b2m_meshify_slab(b2m_ctx *ctx, b2m_comm *comm, const float *d_slab, const int64_t gdims[3], int64_t z0, int64_t nzl,
                     const b2m_opts *opts, b2m_slab_result *out);

/* Same from HOST memory: H2D of this rank's planes, the slab pipeline, D2H of this rank's blocks into malloc()
 * blocks the caller free()s (*verts: nv_edge+nv_cent+nv_extra x 3 f64, *tris: ntris_local x 3 i32). */
int b2m_meshify_slab_host(b2m_ctx *ctx, b2m_comm *comm, const float *h_slab, const int64_t gdims[3], int64_t z0,
                          int64_t nzl, const b2m_opts *opts, void **verts, void **tris, b2m_slab_result *out);

/* ---- atlas front-end: one mesh per label of an indexed volume (replaces the label loop of src/nii2mesh.c:492-583) ----
 * b2m_atlas_scan: nlabel = trunc(max(img)) (:494-501) and, per label i = 0..nlabel, the number of voxels with
 * i-0.5 < img < i+0.5 (:553-563) and their bounding box, from ONE pass over the volume.  *infos is a malloc()'d array
 * of nlabel+1 entries (b2m_atlas_free).  Returns B2M_FAIL (and prints the reference's message) when max(img) < 1.
 * b2m_meshify_label_device: the reference's per-label meshify() - binary volume of that label, isolevel as given
 * (the reference uses 0.5), -l forced off (:493) - computed on the label's bounding box only; the mesh is the one the
 * reference gets from the whole binary volume, bit for bit.  Labels with nvox == 0 are skipped by the reference
 * (:564-567): calling this for one is a usage error. */
typedef struct {
  int label;
  long long nvox;
  int lo[3], hi[3]; /* inclusive bounding box of the label's voxels (undefined when nvox == 0) */
} b2m_label_info;
int b2m_atlas_scan(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], int *nlabel, b2m_label_info **infos);
void b2m_atlas_free(b2m_label_info *infos);
int b2m_meshify_label_device(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const b2m_label_info *info,
                             const b2m_opts *opts, b2m_result *res);

/* The whole label loop in one call (src/nii2mesh.c:540-579, which the reference runs under OpenMP when built with OMP=1):
 * scan, then every non-empty label meshed as b2m_meshify_label_device() does, the labels handed out to `workers` host
 * threads owned by the library (each with its own stream and workspace on ctx's device; <= 0: 8), so that the per-label
 * pipelines overlap on the GPU.  *meshes: malloc()'d array of nlabel+1 entries (index = label; b2m_atlas_meshes_free).
 * rc = 0 meshed, 1 the reference's EXIT_FAILURE for that label, -100 skipped (no voxels), other negative: error.
 * fetch != 0: verts (nverts x 3 f64) / tris (ntris x 3 i32) are copied to malloc()'d host blocks; else counts only. */
typedef struct {
  int label, rc;
  long long nvox;
  int nverts, ntris;
  void *verts, *tris;
  b2m_result r;   /* counts, isolevel actually used, iso_reset, bright box, stage times of this label */
} b2m_label_mesh;
int b2m_atlas_meshify_all(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const b2m_opts *opts, int workers, int fetch,
                          int *nlabel, b2m_label_mesh **meshes);
void b2m_atlas_meshes_free(b2m_label_mesh *meshes, int nlabel);

/* ---- automatic isolevel (-i d / m / b): replaces setThreshold(), src/isolevel.c:245-277 ----
 * dark_medium_bright_123 = 1 dark, 2 medium, 3 bright (src/nii2mesh.c:398-407).  Range, NaN count and the two
 * histograms are GPU reductions; the 256-bin Otsu search runs on the host.  The value equals the reference's float. */
int b2m_isolevel_device(b2m_ctx *ctx, const float *d_img, size_t nvox, int dark_medium_bright_123, float *isolevel);
int b2m_isolevel_host(b2m_ctx *ctx, const float *h_img, size_t nvox, int dark_medium_bright_123, float *isolevel);

/* ---- the formats either side of the path, on the GPU ----
 * b2m_ingest_*: load_nii()'s voxel conversion (src/nii2mesh.c:155-172): raw u8 / i16 / u16 / f32 voxels (NIfTI
 * datatype codes 2 / 4 / 512 / 16) -> f32, (raw * scl_slope) + scl_inter in f32, scl_slope 0 -> 1.  The raw bytes
 * cross PCIe; d_out is a device buffer of nvox floats.
 * b2m_apply_sform_device: apply_sform() (src/meshify.c:1021-1045) on the device mesh of res, in place.
 * b2m_meshify_raw_host: ingest + meshify + (optional) sform + D2H in one call: what nii2() does around meshify(). */
int b2m_ingest_host(b2m_ctx *ctx, const void *h_raw, int datatype, size_t nvox, float scl_slope, float scl_inter, float *d_out);
int b2m_ingest_device(b2m_ctx *ctx, const void *d_raw, int datatype, size_t nvox, float scl_slope, float scl_inter, float *d_out);
int b2m_apply_sform_device(b2m_ctx *ctx, const b2m_result *res, const float srow_x[4], const float srow_y[4],
                           const float srow_z[4]);
int b2m_meshify_raw_host(b2m_ctx *ctx, const void *h_raw, int datatype, const int64_t dims[3], float scl_slope,
                         float scl_inter, const b2m_opts *opts, const float *srow_x, const float *srow_y,
                         const float *srow_z, void **verts, void **tris, b2m_result *res);

/* ---- post-smooth -------------------------------------------------------------------------
 * laplacian_smoothHC() (src/quadric.c:343-394; nii2mesh -s <iter> runs it with alpha 0.1, beta 0.5, lockEdges true,
 * src/nii2mesh.c:331) on a mesh in vec3d / vec3i layout, vertices updated in place; bit-identical to the reference.
 * _device: pointers into device memory (e.g. d_verts / d_tris of a b2m_result); _host: host memory.
 * B2M_EARG when a triangle index lies outside 0..nvert-1 (the reference would read out of bounds). */
int b2m_laplacian_hc_device(b2m_ctx *ctx, double *d_verts, const int *d_tris, int nvert, int ntri, double alpha, double beta,
                            int iter, int lock_edges);
int b2m_laplacian_hc_host(b2m_ctx *ctx, double *h_verts, const int *h_tris, int nvert, int ntri, double alpha, double beta,
                          int iter, int lock_edges);

/* copy the device mesh of the last b2m_meshify_device() call into caller buffers */
int b2m_fetch_mesh(b2m_ctx *ctx, const b2m_result *res, void *h_verts, void *h_tris);

/* ---- stage hooks (parity tests; all pointers are device pointers unless named h_) ---------- */
/* quick_smooth: d_out = smooth(d_in) (out of place). Returns B2M_FAIL and copies in->out if a dim < 5 */
int b2m_stage_smooth(b2m_ctx *ctx, const float *d_in, float *d_out, const int64_t dims[3]);
/* front half of meshify(): smooth, range, sanity, CC mask, fill/largest, darken, bbox.
 * d_composed (optional, N f32): the volume handed to marching cubes (the reference's mutated img).
 * d_mask (optional, N bytes): the bwlabel mask after the optional dilation (mask != 0). */
int b2m_stage_front(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const b2m_opts *opts,
                    float *d_composed, uint8_t *d_mask, b2m_result *res);
/* marching cubes alone on a given (already composed) volume and bbox; mesh left on device in
 * reference emission order (Lewiner) / edge-owner order (classic, welded across cubes). */
int b2m_stage_mc(b2m_ctx *ctx, const float *d_img, const int64_t dims[3], const int lo[3], const int hi[3],
                 const b2m_opts *opts, b2m_result *res);
/* weld + degenerate-triangle removal on a host mesh (uploads, runs the device weld, downloads).
 * h_verts: nv x 3 f64, h_tris: nt x 3 i32, updated in place; *nv,*nt updated. */
int b2m_stage_weld(b2m_ctx *ctx, double *h_verts, int *h_tris, int *nv, int *nt);

#ifdef __cplusplus
}
#endif
#endif /* B2M_H */
