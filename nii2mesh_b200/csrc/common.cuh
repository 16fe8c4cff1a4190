// common.cuh — shared device/host helpers of libb2m (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b2m.h"

#define B2M_WARP 32
#define B2M_NUM_SMS 148 /* B200: 2 dies x 74 SMs; grids of persistent kernels are sized from this */

void b2m_set_error(const char *fmt, ...);

#define CU_TRY(call)                                                                              \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      b2m_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));        \
      return B2M_ECUDA;                                                                           \
    }                                                                                             \
  } while (0)

#define B2M_TRY(call)            \
  do {                           \
    int r_ = (call);             \
    if (r_ != B2M_OK) return r_; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Workspace arena: named, growable device buffers that persist across calls on one ctx so that a
// steady-state meshify() performs no cudaMalloc.
enum {
  BUF_SMOOTH = 0,  // smoothed volume S (N f32)
  BUF_FG,          // foreground bit rows (img >= iso)
  BUF_BG,          // background bit rows (complement, valid x only)
  BUF_FILL,        // fg | bubbles
  BUF_LARGEST,     // largest 18-connected cluster bits
  BUF_KEEP,        // largest | dilate25(largest)
  BUF_CCLIST,      // CC: [count, tile-root slots...]
  BUF_ATLAS,       // atlas: per-label table (scan) / cropped binary sub-volume (one label)
  BUF_MB,          // marching-cubes inside bits (threshold -> composed in place by k_dilate_bbox)
  BUF_NODES,       // union-find nodes (uint2 {parent, size|faceflag}) : 16 per bit word
  BUF_SCALARS,     // small device scalar block (b2m_scalars)
  BUF_SEG,         // MC per-32-voxel segment records (uint4 {xbits,ybits,zbits,vbase})
  BUF_SEG2,        // MC per-segment {tbase, cbase}
  BUF_ACTIVE,      // MC active-voxel records (uint4)
  BUF_VERTS,       // f64 xyz
  BUF_TRIS,        // i32 xyz
  BUF_VERTS2,      // compacted vertices
  BUF_TRIS2,       // compacted triangles
  BUF_CAND,        // weld candidates (u32 vertex ids)
  BUF_SORTA,       // sort ping
  BUF_SORTB,       // sort pong
  BUF_SORTH,       // sort histograms
  BUF_REMAP,       // u32 per vertex
  BUF_FLAGS,       // u32 per vertex / triangle (keep flags -> scanned)
  BUF_SCAN1,       // scan partials
  BUF_SCAN2,
  BUF_TABLES,      // MC case tables blob
  BUF_L2FLUSH,
  BUF_TMP0,
  BUF_TMP1,
  BUF_INPUT,       // device copy of a host volume (b2m_meshify_host)
  BUF_WELD,        // weld: all the small per-item arrays and block tables (carved)
  BUF_ITEMS_ALL,   // slabs: weld items of all ranks
  BUF_HALO_LO,     // slabs: raw planes received from the rank below (smooth halo)
  BUF_HALO_HI,     // slabs: raw planes received from the rank above
  BUF_HALO_V,      // slabs: the next rank's first-plane vertices (triangle pass)
  BUF_SEAM0,       // slabs: CC seam lists (roots, dense ids, pairs, tables) carved per use
  BUF_SEAM1,
  BUF_SEAM2,
  BUF_SEAMF,       // slabs: fast seam path (gathered blocks, replicated union-find, dense ticket planes)
  BUF_GATHER,      // slabs: padded segments of a variable-length all-gather
  BUF_POST_INC,    // post-smooth: vertex -> (triangle, corner) incidence list
  BUF_POST_P,      // post-smooth: p / p' / b vertex arrays + border flags
  BUF_POST_V,      // post-smooth from host memory: device copy of the vertices
  BUF_POST_T,      //   ... and of the triangles
  BUF_COUNT
};

struct b2m_buf {
  void *p;
  size_t cap;
};

// device-side scalars of one meshify call (one 256-byte block, zeroed per call).  In slab mode the
// blocks of all ranks are all-gathered at every host decision point (b2m_sync_scalars).
struct b2m_scalars {
  unsigned int vmin_enc, vmax_enc;       // order-preserving encodings of f32 min/max
  unsigned int cmin_enc;                 // min of the composed volume (lazy)
  int lo[3], hi[3];                      // bright bbox (raw, before the +-1/+2 widening), global coordinates
  unsigned long long best_fg;            // (size << 32) | ~rootslot  of the largest fg cluster (of this rank)
  unsigned int nroots_fg, nroots_bg;     // number of components (of this rank)
  unsigned int n_active;                 // MC active records appended
  unsigned int n_cand;                   // weld items appended
  unsigned int n_first;                  // slabs: edge vertices owned by the first own sub-volume plane
  unsigned int tot_v, tot_t, tot_c;      // MC totals (edge vertices, triangles, centroid vertices)
  unsigned int n_tri_kept;               // triangles surviving the degenerate test
  unsigned int overflow;                 // capacity overflow / consistency flags
  unsigned int seam_n;                   // slabs: seam list length (roots, then unique roots, then pairs)
  unsigned int seam_roots;               // slabs: components of the replicated seam union-find
  unsigned long long first_cube;         // min (row << 16 | x) over active cubes (classic pts[0])
  double pts0[3];                        // classic: first soup vertex (the weld's key origin)
  double v0[3];                          // slabs: this rank's first vertex (Lewiner key origin = global vertex 0)
  unsigned long long best_seam;          // slabs: largest seam component - slow path (size << 30) | (2^30-1 - seam id), fast path its size
  unsigned long long best_seam_prio;     // slabs, fast path: (rank << 32 | run slot) of that component's first voxel
  // connected components of >= 2^31 voxels: a node's count field holds 31 bits; every wrap is recorded here
  // (at most two such components fit a slab of <= 2^32 voxels).  Reset before each labelling.
  unsigned int carry_slot[4];            // root slot + 1 (0: free)
  unsigned int carry_hi[4];              // wraps = units of 2^31 voxels
  // weld bases of this rank (k_w_bases), read with the last host sync of the call instead of a sync of their own
  unsigned int wb_new_e_off, wb_nve_new, wb_new_c_base, wb_nvc_new, wb_n_dead, wb_n_extra;
  unsigned int pad[7];
  unsigned int sync_seq;                 // LAST word of the block: stamped on the device before the block is copied to the
                                         // host, so that the host can wait for the copy by watching pinned memory
};
static_assert(sizeof(b2m_scalars) == 256, "b2m_scalars is exchanged as one 256-byte block");

// one weld item: a vertex (Lewiner) or a soup copy (classic) close enough to a grid corner to merge
struct b2m_item {
  double pos[3];
  uint32_t id;   // original index in the reference's vertex array (vertex id / soup index 3*t+c)
  uint32_t vid;  // edge-keyed vertex id it belongs to
};

// ---------------------------------------------------------------------------------------------
// z-slab of a volume owned by one rank.  EXT buffers (S, bit rows) hold the own planes plus one halo
// plane per existing neighbour: ext plane e <-> global plane e + ez0.
struct b2m_slab {
  int rank, world;
  int gnz;       // global NZ
  int z0, nzl;   // own planes: global [z0, z0 + nzl)
  int hl, hh;    // halo planes below / above (0 or 1)
  int ez0, nze;  // ext origin (z0 - hl) and ext plane count (hl + nzl + hh)
};
struct b2m_comm;
int b2m_comm_rank(const b2m_comm *c);
int b2m_comm_world(const b2m_comm *c);
void b2m_comm_abort(b2m_comm *c);
// all-gather the scalar blocks of all ranks (device -> pinned host) and synchronise the stream
int b2m_sync_scalars(b2m_ctx *ctx, b2m_comm *c);
b2m_scalars *b2m_sc(b2m_ctx *ctx, b2m_comm *c, int rank);  // host copy of rank's block after b2m_sync_scalars
// neighbour exchange along z (byte counts; 0 = nothing in that direction; ignored at the ends)
int b2m_comm_exchange(b2m_ctx *ctx, b2m_comm *c, const void *d_send_up, size_t send_up_bytes, void *d_recv_lo,
                      size_t recv_lo_bytes, const void *d_send_dn, size_t send_dn_bytes, void *d_recv_hi,
                      size_t recv_hi_bytes);
int b2m_comm_exchange_on(b2m_ctx *ctx, b2m_comm *c, cudaStream_t st, const void *d_send_up, size_t send_up_bytes, void *d_recv_lo,
                         size_t recv_lo_bytes, const void *d_send_dn, size_t send_dn_bytes, void *d_recv_hi, size_t recv_hi_bytes);
bool b2m_comm_async_capable(const b2m_comm *c);
int b2m_comm_stream_wait(b2m_ctx *ctx, b2m_comm *c);  // host wait for the ctx stream that a failed peer can interrupt
int b2m_comm_allgather_inplace(b2m_ctx *ctx, b2m_comm *c, void *d_buf, size_t bytes);  // rank r's block at d_buf + r * bytes
// d_recv = concatenation over ranks of their d_send (bytes[r] each); every rank passes the same bytes[]
int b2m_comm_allgatherv(b2m_ctx *ctx, b2m_comm *c, const void *d_send, void *d_recv, const size_t *bytes);
int b2m_comm_gather_items(b2m_ctx *ctx, b2m_comm *c, unsigned n_local, b2m_item **items, unsigned *n);

// per-kernel device timing (opt-in, b2m_set_profile): one CUDA-event pair per launch on ctx->stream
#define B2M_KT_MAX 256
struct b2m_ktimer {
  const char *name;
  cudaEvent_t e0, e1;
};

#define B2M_ATLAS_WORKERS_MAX 32
#define B2M_PEND_MAX 16
#define B2M_RING_CHUNK ((size_t)8 << 20)   /* streamed D2H: DMA granule */
#define B2M_RING_SLOTS 12                  /* = 3 * B2M_STAGE_BYTES / B2M_RING_CHUNK */
struct b2m_ctx {
  int device;
  b2m_scalars *h_all;      // pinned: scalar blocks of all ranks (slabs), h_all_cap entries
  b2m_scalars *d_all;
  int h_all_cap;
  unsigned slab_t_off;     // slabs: global index of this rank's first surviving triangle (last call)
  // called once per hot-path call when the marching-cubes totals are known (own vertices / triangles before the
  // weld): the host entry points use it to allocate and pre-fault the output blocks while the GPU keeps working
  void (*counts_hook)(void *user, size_t nverts, size_t ntris);
  void *hook_user;
  int origin[3];           // atlas: where the (cropped) volume of the current call sits in the caller's volume
  size_t last_nvox, last_nv, last_nt;  // geometry and marching-cubes totals of the previous host call
  cudaStream_t stream;
  b2m_buf buf[BUF_COUNT];
  cudaEvent_t ev[2 * B2M_NSTAGE + 2];
  b2m_scalars *h_scalars;  // pinned mirror
  b2m_scalars *h_land;     // pinned landing block of b2m_sync_scalars (watched for sync_seq)
  unsigned sync_seq;
  uint64_t launches;
  int tables_ready;
  int smooth_tma_attr_done;
  int smooth_attr_done;    // the > 48 KB dynamic shared memory opt-in of the smooth kernel was made through this ctx
  int sm_count;
  unsigned ev_mask;        // which stage event pairs were recorded in the current call
  // host<->device staging for pageable host memory (b2m_copy_h2d / b2m_copy_d2h)
  void *stage[3];          // pinned ring buffers (B2M_STAGE_BYTES each), allocated on first use
  cudaEvent_t stage_ev[3];
  cudaEvent_t ring_ev[B2M_RING_SLOTS];  // streamed D2H: the three buffers seen as B2M_RING_SLOTS sub-chunks
  // H2D overlapped with the smooth (B2M_H2D_OVERLAP=1, experimental): the volume arrives in z-chunks on copy_stream;
  // pend_zend[k] raw planes are on the device once pend_ev[k] has fired.  Consumed (and cleared) by b2m_front_run
  cudaStream_t copy_stream;
  cudaEvent_t pend_ev[B2M_PEND_MAX + 1];  // [B2M_PEND_MAX] = start of the transfer
  int pend_zend[B2M_PEND_MAX];
  int pend_n, pend_last;   // chunks not yet consumed / index of the last chunk of the transfer
  cudaStream_t aux_stream; // slabs: halo exchanges that run next to kernels of `stream` (high priority, created on demand)
  cudaEvent_t aux_ev[2];
  b2m_ctx *atlas_workers[B2M_ATLAS_WORKERS_MAX];  // worker contexts of b2m_atlas_meshify_all (created on demand, same device)
  int profile;             // record an event pair around every kernel launch
  int nkt, nkt_events;     // entries used in this call / event pairs created so far
  b2m_ktimer kt[B2M_KT_MAX];
};

int b2m_kt_begin(b2m_ctx *ctx, const char *name);  // returns slot or -1
void b2m_kt_end(b2m_ctx *ctx, int slot);
#define KT_LAUNCH(ctx, name, ...)             \
  do {                                        \
    int kt_ = b2m_kt_begin((ctx), (name));    \
    __VA_ARGS__;                              \
    b2m_kt_end((ctx), kt_);                   \
    (ctx)->launches++;                        \
  } while (0)

int b2m_reserve(b2m_ctx *ctx, int which, size_t bytes);
template <typename T>
static inline T *b2m_ptr(b2m_ctx *ctx, int which) {
  return reinterpret_cast<T *>(ctx->buf[which].p);
}
int b2m_aux_stream(b2m_ctx *ctx);     // creates ctx->aux_stream / aux_ev on first use
int b2m_fetch_scalars(b2m_ctx *ctx);  // D2H of the scalar block + stream sync
// bulk copies between device memory and ANY host memory (pinned: one DMA; pageable: pipelined through
// pinned ring buffers with a multi-threaded host memcpy); synchronous on return
#define B2M_STAGE_BYTES ((size_t)32 << 20)
int b2m_copy_h2d(b2m_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int b2m_copy_d2h(b2m_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
bool b2m_host_is_pinned(const void *p);
// start an asynchronous z-chunked H2D of a pinned volume; b2m_front_run waits chunk by chunk (ctx->pend_*)
int b2m_h2d_chunked_begin(b2m_ctx *ctx, float *d_dst, const float *h_src, size_t nxy, int nz);
int b2m_h2d_chunked_ms(b2m_ctx *ctx, float *ms);  // duration of the transfer started by the call above
bool b2m_d2h_registers(void);  // the D2H of big pageable blocks registers the destination piecewise (few-threads mode)
int b2m_copy_d2h_f32exact(b2m_ctx *ctx, double *h_dst, const double *d_src, size_t n, int *done);  // doubles that are all f32 values: 4 B each over PCIe
int b2m_touch_async(void *a, size_t na, void *b, size_t nb);
void b2m_touch_wait(void);

#define B2M_LAUNCHED(ctx) ((ctx)->launches++)

static inline unsigned int b2m_cdiv(size_t a, size_t b) { return (unsigned int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// order-preserving f32 <-> u32 (for atomicMin/atomicMax on floats)
__host__ __device__ static inline unsigned int f32_enc(float f) {
  unsigned int u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ static inline float f32_dec(unsigned int e) {
  unsigned int u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  float f;
#ifdef __CUDA_ARCH__
  f = __uint_as_float(u);
#else
  memcpy(&f, &u, 4);
#endif
  return f;
}

// ---------------------------------------------------------------------------------------------
// volume geometry shared by the kernels
struct b2m_geom {
  int nx, ny, nz;      // volume dims (slabs: the EXT planes of this rank)
  int w;               // 32-bit words per bit row = ceil(nx/32)
  long long nxy;       // nx*ny
  long long n;         // voxels
  long long nrows;     // ny*nz
  long long nwords;    // nrows*w
};
static inline b2m_geom b2m_make_geom(const int64_t dims[3]) {
  b2m_geom g;
  g.nx = (int)dims[0]; g.ny = (int)dims[1]; g.nz = (int)dims[2];
  g.w = (g.nx + 31) / 32;
  g.nxy = (long long)g.nx * g.ny;
  g.n = g.nxy * g.nz;
  g.nrows = (long long)g.ny * g.nz;
  g.nwords = g.nrows * g.w;
  return g;
}

// ---------------------------------------------------------------------------------------------
// composed value of one voxel = what the reference's mutated img holds when marching cubes runs
// (/root/reference/src/meshify.c:332-365): bubble fill to >= iso, non-kept voxels to mn, faces
// darkened to <= edge_max.  Never materialised on the hot path: S plus two bit rows.
// Slabs: S / fill / keep are the EXT buffers pre-offset by -ez0 planes, so kernels index them with
// GLOBAL z (only planes the rank holds are ever touched); nz is the global NZ.
struct compose_params {
  const float *S;
  const uint32_t *fill;
  const uint32_t *keep;
  int nx, ny, nz, w;
  float iso, mn, edge_max;
};
__device__ __forceinline__ float composed_value(const compose_params &c, int x, int y, int z) {
  size_t row = (size_t)z * c.ny + y;
  float v = __ldg(c.S + row * c.nx + x);
  if (c.fill || c.keep) {
    size_t word = row * c.w + (x >> 5);
    if (c.fill && ((__ldg(c.fill + word) >> (x & 31)) & 1u)) v = fmaxf(v, c.iso);
    if (c.keep && !((__ldg(c.keep + word) >> (x & 31)) & 1u)) v = c.mn;
  }
  if (x == 0 || y == 0 || z == 0 || x == c.nx - 1 || y == c.ny - 1 || z == c.nz - 1) v = fminf(c.edge_max, v);
  return v;
}

// stages implemented across the .cu files ------------------------------------------------------
// raw input of the smooth as seen by one rank: own planes plus the halo planes received from the neighbours
struct smooth_src {
  const float *lo, *main, *hi;
  int n_lo, n_main, n_hi;  // planes in each piece; raw plane q = global z - rz0
  int rz0;                 // global z of raw plane 0
  int gnz;                 // global NZ
  int oz0, onz;            // output: global planes [oz0, oz0+onz) -> d_out plane (z - oz0)
};
// optional by-product of the smooth: the threshold bit rows of its output planes for a GIVEN isolevel (see k_smooth3).
// The pointers are offset to output plane 0 of the call, like d_out; bg may be null.
struct smooth_bits {
  uint32_t *fg, *bg, *mb;
  float iso;
  int classic;
};
// *bits_done = 1 when the call wrote the bit rows (vector path, nx % 32 == 0, default kernel), else 0
int b2m_smooth_run(b2m_ctx *ctx, const smooth_src &src, float *d_out, const b2m_geom &g, b2m_scalars *d_sc,
                   const smooth_bits *bits = nullptr, int *bits_done = nullptr);
int b2m_minmax_run(b2m_ctx *ctx, const float *d_in, size_t n, b2m_scalars *d_sc);
// fg = v >= iso, bg = its complement (optional), mb = the marching-cubes comparison (optional):
// Lewiner v - iso > -FLT_EPSILON (src/MarchingCubes.c:132-133 after :1112), classic v < iso (src/oldcubes.c:407-414)
int b2m_threshold_run(b2m_ctx *ctx, const float *d_in, const b2m_geom &g, float iso, uint32_t *d_fg, uint32_t *d_bg,
                      uint32_t *d_mb, int classic);

struct b2m_front_out {
  const float *S;          // smoothed (or original) volume, EXT layout
  const uint32_t *fill;    // nullptr or fg|bubbles bit rows, EXT layout
  const uint32_t *keep;    // nullptr or largest|dilated bit rows, EXT layout
  const uint32_t *ibits;   // marching-cubes inside bit of every voxel of the composed volume, EXT layout
  float iso, vmin, vmax, edge_max;
  int lo[3], hi[3];        // widened bbox as handed to marching cubes (global)
  int iso_reset;
  int bits_ready;          // the smooth already wrote fg / bg / mb for isolevel bits_iso (b2m_cc_run then skips k_threshold)
  float bits_iso;
};
int b2m_cc_run(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const b2m_geom &g, const b2m_opts *o, b2m_scalars *d_sc,
               b2m_front_out *fo);
int b2m_front_run(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const float *d_img, const b2m_geom &g, const b2m_opts *o,
                  b2m_front_out *fo, b2m_result *res);
int b2m_compose_materialize(b2m_ctx *ctx, const b2m_geom &g, const b2m_front_out *fo, float *d_composed,
                            uint8_t *d_mask, b2m_scalars *d_sc, int want_min);

// marching-cubes output of one rank, before the weld.  Vertex ids are GLOBAL (reference emission order
// over the whole volume): own edge vertices e_off .. e_off+nv_edge, own centroid vertices
// NVE+c_off .. NVE+c_off+nv_c; verts[] holds the own blocks back to back.
// A triangle whose three vertices are edge vertices farther than d from the ends of their cube edges has area >= 0.3 d^2
// (all sides >= d, every height >= d/2: two points on different cube edges are at least max(d_a, d_b) apart, and a
// boundary point near the line through two others must be near one of them or near a corner), so with d = 1/128 it is
// three orders of magnitude above the reference's FLT_EPSILON area threshold (src/meshify.c:122-142): only triangles
// with a flagged or centroid vertex need the FP64 needle test and its three position gathers.
#define B2M_NEAR_TOL 0.0078125f
struct b2m_mesh_dev {
  double *verts;  // BUF_VERTS: [edge block | centroid block]
  int *tris;      // BUF_TRIS (global vertex ids)
  unsigned int nv_edge, nv_c, nt;   // own counts
  unsigned int NVE, NVC, NT;        // global counts
  unsigned int e_off, c_off, t_off; // global offsets of the own blocks
  unsigned int nitems;              // own weld items in BUF_CAND
  const uint32_t *nearbits;         // one bit per own EDGE vertex (local index): within B2M_NEAR_TOL of a grid corner
  const double *halo_verts;         // next rank's first-plane vertices: ids halo0 .. halo1
  unsigned int halo0, halo1;
  const double *d_p0;               // device pointer to the weld's key origin (the reference's pts[0])
  int classic_soup;                 // items are soup copies (classic back-end)
};
struct b2m_weld_out {
  const double *verts;  // own welded vertices: [edge | centroid | extras]
  const int *tris;      // own surviving triangles, welded global vertex indices
  unsigned int nv_local, nve_local, nvc_local, nx_local, nt_local;
  unsigned int v_edge_off, v_c_off;  // global welded index of the first own edge / centroid vertex
  unsigned int nv_global, n_dead, n_extra;
  unsigned int n_items;  // weld items the run worked on (all ranks')
};
int b2m_mc_run(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const b2m_geom &g, const b2m_opts *o,
               const b2m_front_out *fo, b2m_mesh_dev *mesh);
int b2m_weld_run(b2m_ctx *ctx, b2m_comm *comm, b2m_mesh_dev *mesh, int all_items, b2m_weld_out *wo);   // enqueue only
int b2m_weld_finish(b2m_ctx *ctx, b2m_comm *comm, const b2m_mesh_dev *mesh, b2m_weld_out *wo);        // after b2m_sync_scalars

// generic primitives (scan.cu)
int b2m_exclusive_scan_u32(b2m_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *d_total);
int b2m_sort_u64(b2m_ctx *ctx, uint64_t *d_keys, size_t n, int key_bits);  // ascending, in place (uses BUF_SORTB/H)
