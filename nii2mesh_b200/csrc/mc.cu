// mc.cu — kernel 3: marching cubes as classify -> prefix scan -> compact emit.
//
// Three reference flavours behind one runtime switch (the reference picks at compile time,
// /root/reference/src/meshify.c:25-29):
//   backend LEWINER, original_mc = 0 : MC33 with face/interior ambiguity tests
//        src/MarchingCubes.c:116-155 (cube loop), :235-270 (edge vertices), :276-295 (test_face),
//        :301-453 (test_interior), :458-795 (process_cube), :803-859 (add_triangle),
//        :935-1026 (edge vertex = (float)i + c0/(c0-c1)), :1029-1082 (centroid vertex),
//        :1086-1143 (sub-volume copy with -isolevel, export with +lo in f32, reversed winding)
//   backend LEWINER, original_mc = 1 : same code, `casesClassic` table (:471-477)
//   backend CLASSIC                  : src/oldcubes.c:465-522 / :50-463 / :22-40 — FP64 positions
//        p1 + mu*(p2-p1); the reference emits a triangle SOUP which its weld then merges; here
//        vertices are born welded per grid edge (edge-keyed), taking the FP64 variant of the soup
//        copy with the highest soup index, which is the one the reference's weld keeps
//        (src/meshify.c:99-100).
//
// The composed volume (bubble fill, largest-cluster mask, darkened faces) is never materialised:
// every load goes through composed_value() (S + two bit rows), see cc.cu.
//
// Output order = the reference's emission order: edge vertices in raster order of their owner
// voxel (x, then y, then z edge per voxel), then centroid vertices in cube raster order;
// triangles in cube raster order.  That is obtained from an exclusive scan over per-32-voxel
// segment counts; the emit pass runs only over the compacted list of active voxels.
#include "common.cuh"

#define MCT_NO_BLOB
#include "mc_tables.inc"
#undef MCT_NO_BLOB


struct mc_params {
  compose_params c;
  int lo0, lo1, lo2;
  int org0, org1, org2;  // atlas: position of this (cropped) volume inside the caller's volume; exported coordinates only
  int sx, sy, sz;   // sub-volume size in voxels
  int zs0, zn;      // slabs: sub-volume planes [zs0, zs0+zn) are owned by this rank (segment arrays hold zn+1 planes:
                    // the last one is the next rank's first plane, received after the scan)
  int segs;         // 32-voxel segments per sub-volume row
  float pad;        // Lewiner: value of sub-volume voxels that fall outside the volume (already - iso)
  unsigned pad_inside;  // its inside bit
  int closed;           // slabs: faces are strictly outside, so every wrapped voxel is outside too and is never read
                        // (its plane may not be on this rank)
  int classic;      // backend == CLASSIC
  int original_mc;
  const signed char *tab;
  const uint32_t *lutinfo;  // 256 entries: Lewiner MC33, or the classic table when classic / originalMC
  const uint32_t *ibits;    // inside bit per voxel of the composed volume (bit rows, indexed with global z)
  uint4 *segbits;   // per segment: x/y/z edge-vertex bit masks, .w = packed counts -> vertex base (scan3)
  uint32_t *segt, *segc;  // per segment exclusive triangle / centroid-vertex bases
  uint32_t *segcnt;       // per segment packed counts (dense copy for the scan: 4 B instead of a 16-byte stride)
  uint4 *active;
  unsigned int active_cap;
  b2m_scalars *sc;
};

// value of sub-volume voxel (x,y,z).  Lewiner: img[j]-iso with the reference's linear-index-only
// guard (row wrap quirk, src/MarchingCubes.c:1106-1115), clamped away from zero (:132,:252-255).
// Classic: the raw composed value.
__device__ __forceinline__ float mc_data(const mc_params &p, int x, int y, int z) {
  int gx = p.lo0 + x, gy = p.lo1 + y, gz = p.lo2 + z;
  if (p.classic) return composed_value(p.c, gx, gy, gz);
  if (gx >= p.c.nx) { gx -= p.c.nx; gy++; }
  if (gy >= p.c.ny) { gy -= p.c.ny; gz++; }
  float v = (gz >= p.c.nz) ? p.pad : composed_value(p.c, gx, gy, gz) - p.c.iso;
  if (fabsf(v) < FLT_EPSILON) v = FLT_EPSILON;
  return v;
}
__device__ __forceinline__ bool mc_inside(const mc_params &p, float v) { return p.classic ? (v < p.c.iso) : (v > 0.0f); }

// ---- MC33 decision procedure ---------------------------------------------------------------
__device__ const signed char MC_FACE_Q[6][4] = {{0, 4, 5, 1}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 4, 0}, {0, 3, 2, 1}, {4, 7, 6, 5}};
// rows {e0,e1, B0,B1, C0,C1, D0,D1}: the twelve reference-edge slices of test_interior (:335-420)
__device__ const signed char MC_EDGE_SLICE[12][8] = {
    {0, 1, 3, 2, 7, 6, 4, 5}, {1, 2, 0, 3, 4, 7, 5, 6}, {2, 3, 1, 0, 5, 4, 6, 7}, {3, 0, 2, 1, 6, 5, 7, 4},
    {4, 5, 7, 6, 3, 2, 0, 1}, {5, 6, 4, 7, 0, 3, 1, 2}, {6, 7, 5, 4, 1, 0, 2, 3}, {7, 4, 6, 5, 2, 1, 3, 0},
    {0, 4, 3, 7, 2, 6, 1, 5}, {1, 5, 0, 4, 3, 7, 2, 6}, {2, 6, 1, 5, 0, 4, 3, 7}, {3, 7, 2, 6, 1, 5, 0, 4}};

__device__ const signed char MC_EDGE_A[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
__device__ const signed char MC_EDGE_B[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};

__device__ __forceinline__ bool t_face(const float *c, int f) {  // src/MarchingCubes.c:276-295
  const signed char *q = MC_FACE_Q[(f < 0 ? -f : f) - 1];
  float A = c[q[0]], B = c[q[1]], C = c[q[2]], D = c[q[3]];
  float det = __fsub_rn(__fmul_rn(A, C), __fmul_rn(B, D));
  if (fabsf(det) < FLT_EPSILON) return f >= 0;
  return __fmul_rn(__fmul_rn((float)f, A), det) >= 0.0f;
}

__device__ __forceinline__ float lerp_rn(float a, float b, float t) {  // a + (b - a) * t, no contraction
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

__device__ bool t_interior(const float *c, int kase, int refedge, int s) {  // :301-453
  float t, At = 0.f, Bt = 0.f, Ct = 0.f, Dt = 0.f;
  if (kase == 4 || kase == 10) {
    float d40 = __fsub_rn(c[4], c[0]), d62 = __fsub_rn(c[6], c[2]), d73 = __fsub_rn(c[7], c[3]), d51 = __fsub_rn(c[5], c[1]);
    float a = __fsub_rn(__fmul_rn(d40, d62), __fmul_rn(d73, d51));
    float b = __fmul_rn(c[2], d40);
    b = __fadd_rn(b, __fmul_rn(c[0], d62));
    b = __fsub_rn(b, __fmul_rn(c[1], d73));
    b = __fsub_rn(b, __fmul_rn(c[3], d51));
    t = __fdiv_rn(-b, __fmul_rn(2.0f, a));
    if (t < 0.f || t > 1.f) return s > 0;
    At = __fadd_rn(c[0], __fmul_rn(d40, t));
    Bt = __fadd_rn(c[3], __fmul_rn(d73, t));
    Ct = __fadd_rn(c[2], __fmul_rn(d62, t));
    Dt = __fadd_rn(c[1], __fmul_rn(d51, t));
  } else if (refedge >= 0 && refedge < 12) {
    const signed char *r = MC_EDGE_SLICE[refedge];
    t = __fdiv_rn(c[r[0]], __fsub_rn(c[r[0]], c[r[1]]));
    At = 0.f;
    Bt = lerp_rn(c[r[2]], c[r[3]], t);
    Ct = lerp_rn(c[r[4]], c[r[5]], t);
    Dt = lerp_rn(c[r[6]], c[r[7]], t);
  }
  int test = (At >= 0.f ? 1 : 0) | (Bt >= 0.f ? 2 : 0) | (Ct >= 0.f ? 4 : 0) | (Dt >= 0.f ? 8 : 0);
  float det = __fsub_rn(__fmul_rn(At, Ct), __fmul_rn(Bt, Dt));
  switch (test) {
    case 5: return (det < FLT_EPSILON) ? s > 0 : s < 0;
    case 10: return (det >= FLT_EPSILON) ? s > 0 : s < 0;
    case 7: case 11: case 13: case 14: case 15: return s < 0;
    default: return s > 0;
  }
}

#define TROW(name, cfg) (MCT_##name + (cfg) * MCT_##name##_ROW)
#define TSUB(name, cfg, sub) (MCT_##name + (cfg) * MCT_##name##_ROW + (sub) * MCT_##name##_SUB)
#define PICK(off_, n_, c_) do { off = (off_); ntri = (n_); hasc = (c_); return; } while (0)

// src/MarchingCubes.c:458-795.  c = clamped corner values; lut bit p = c[p] > 0.
__device__ __noinline__ void mc33_select(const signed char *__restrict__ tab, const float *c, int lut, int original_mc, int &off,
                            int &ntri, int &hasc) {
  if (original_mc) {
    int o = MCT_casesClassic + 16 * lut, n = 0;
    while (n < 5 && tab[o + 3 * n] != -1) n++;
    PICK(o, n, 0);
  }
  const int kase = tab[MCT_cases + 2 * lut], cfg = tab[MCT_cases + 2 * lut + 1];
  switch (kase) {
    case 1: PICK(TROW(tiling1, cfg), 1, 0);
    case 2: PICK(TROW(tiling2, cfg), 2, 0);
    case 3:
      if (t_face(c, tab[MCT_test3 + cfg])) PICK(TROW(tiling3_2, cfg), 4, 0);
      PICK(TROW(tiling3_1, cfg), 2, 0);
    case 4:
      if (t_interior(c, 4, -1, tab[MCT_test4 + cfg])) PICK(TROW(tiling4_1, cfg), 2, 0);
      PICK(TROW(tiling4_2, cfg), 6, 0);
    case 5: PICK(TROW(tiling5, cfg), 3, 0);
    case 6: {
      const signed char *t = tab + TROW(test6, cfg);
      if (t_face(c, t[0])) PICK(TROW(tiling6_2, cfg), 5, 0);
      if (t_interior(c, 6, t[2], t[1])) PICK(TROW(tiling6_1_1, cfg), 3, 0);
      PICK(TROW(tiling6_1_2, cfg), 9, 1);
    }
    case 7: {
      const signed char *t = tab + TROW(test7, cfg);
      int sub = (t_face(c, t[0]) ? 1 : 0) | (t_face(c, t[1]) ? 2 : 0) | (t_face(c, t[2]) ? 4 : 0);
      switch (sub) {
        case 0: PICK(TROW(tiling7_1, cfg), 3, 0);
        case 1: PICK(TSUB(tiling7_2, cfg, 0), 5, 0);
        case 2: PICK(TSUB(tiling7_2, cfg, 1), 5, 0);
        case 3: PICK(TSUB(tiling7_3, cfg, 0), 9, 1);
        case 4: PICK(TSUB(tiling7_2, cfg, 2), 5, 0);
        case 5: PICK(TSUB(tiling7_3, cfg, 1), 9, 1);
        case 6: PICK(TSUB(tiling7_3, cfg, 2), 9, 1);
        default:
          if (t_interior(c, 7, t[4], t[3])) PICK(TROW(tiling7_4_2, cfg), 9, 0);
          PICK(TROW(tiling7_4_1, cfg), 5, 0);
      }
    }
    case 8: PICK(TROW(tiling8, cfg), 2, 0);
    case 9: PICK(TROW(tiling9, cfg), 4, 0);
    case 10: {
      const signed char *t = tab + TROW(test10, cfg);
      bool f0 = t_face(c, t[0]), f1 = t_face(c, t[1]);
      if (f0 && f1) PICK(TROW(tiling10_1_1_, cfg), 4, 0);
      if (f0) PICK(TROW(tiling10_2, cfg), 8, 1);
      if (f1) PICK(TROW(tiling10_2_, cfg), 8, 1);
      if (t_interior(c, 10, -1, t[2])) PICK(TROW(tiling10_1_1, cfg), 4, 0);
      PICK(TROW(tiling10_1_2, cfg), 8, 0);
    }
    case 11: PICK(TROW(tiling11, cfg), 4, 0);
    case 12: {
      const signed char *t = tab + TROW(test12, cfg);
      bool f0 = t_face(c, t[0]), f1 = t_face(c, t[1]);
      if (f0 && f1) PICK(TROW(tiling12_1_1_, cfg), 4, 0);
      if (f0) PICK(TROW(tiling12_2, cfg), 8, 1);
      if (f1) PICK(TROW(tiling12_2_, cfg), 8, 1);
      if (t_interior(c, 12, t[3], t[2])) PICK(TROW(tiling12_1_1, cfg), 4, 0);
      PICK(TROW(tiling12_1_2, cfg), 8, 0);
    }
    case 13: {
      const signed char *t = tab + TROW(test13, cfg);
      int sub = 0;
#pragma unroll
      for (int b = 0; b < 6; b++)
        if (t_face(c, t[b])) sub |= 1 << b;
      const int sc = tab[MCT_subconfig13 + sub];
      if (sc == 0) PICK(TROW(tiling13_1, cfg), 4, 0);
      if (sc >= 1 && sc <= 6) PICK(TSUB(tiling13_2, cfg, sc - 1), 6, 0);
      if (sc >= 7 && sc <= 18) PICK(TSUB(tiling13_3, cfg, sc - 7), 10, 1);
      if (sc >= 19 && sc <= 22) PICK(TSUB(tiling13_4, cfg, sc - 19), 12, 1);
      if (sc >= 23 && sc <= 26) {
        const int k = sc - 23;
        const int refedge = tab[TSUB(tiling13_5_1, cfg, k)];
        if (t_interior(c, 13, refedge, t[6])) PICK(TSUB(tiling13_5_1, cfg, k), 6, 0);
        PICK(TSUB(tiling13_5_2, cfg, k), 10, 0);
      }
      if (sc >= 27 && sc <= 38) PICK(TSUB(tiling13_3_, cfg, sc - 27), 10, 1);
      if (sc >= 39 && sc <= 44) PICK(TSUB(tiling13_2_, cfg, sc - 39), 6, 0);
      if (sc == 45) PICK(TROW(tiling13_1_, cfg), 4, 0);
      PICK(0, 0, 0);  // "Impossible case 13?" (:785): no triangles
    }
    case 14: PICK(TROW(tiling14, cfg), 4, 0);
    default: PICK(0, 0, 0);
  }
}

// ---- pass A: classify ------------------------------------------------------------------------
// Works on BITS, 32 cubes per thread.  The front end leaves one "inside" bit per voxel of the composed volume
// (p.ibits: threshold at the marching-cubes comparison, bubble fill, largest-cluster mask and face darkening
// folded in at word level, cc.cu:k_dilate_bbox), so a thread classifies a 32-voxel segment of a sub-volume row
// with a few word operations: edge-vertex masks are XORs of neighbouring rows, the active-cube mask is
// OR & ~AND of the eight corner words.  A thread owns one (row y, segment) column and marches along z, keeping
// the two rows of the current plane in registers.  Only ACTIVE voxels (~5 %) cost per-voxel work: cube index
// from the eight words -> per-lut summary table; corner VALUES are read only for the ambiguous MC33 cases.
// Per segment the kernel leaves {xbits, ybits, zbits, packed counts}; packed counts = nv | nt << 7 | nc << 16
// (nv <= 96, nt <= 384, nc <= 32), turned into exclusive bases by scan3.  Active voxels are appended to a
// warp-private shared-memory buffer and flushed to the global list with one atomic per ~150 records.
#define MCB_THREADS 128
#define MCB_ZC 8
#define MCB_BUF 192

// inside bit of one sub-volume voxel given in VOLUME coordinates that may lie one column / row / plane beyond
// the volume: the reference's linear-index arithmetic wraps the extra column into the next row and the extra row
// into the next plane; only the extra plane is padding (src/MarchingCubes.c:1106-1115)
__device__ __forceinline__ uint32_t mc_inside_bit_wrapped(const mc_params &p, int gx, int gy, int gz) {
  if (gx >= p.c.nx) { if (p.closed) return 0u; gx -= p.c.nx; gy++; }
  if (gy >= p.c.ny) { if (p.closed) return 0u; gy -= p.c.ny; gz++; }
  if (gz >= p.c.nz) return p.pad_inside;
  return (__ldg(p.ibits + ((size_t)gz * p.c.ny + gy) * p.c.w + (gx >> 5)) >> (gx & 31)) & 1u;
}

// the 33 inside bits of sub-volume row (y, z) that start at sub-volume x = seg*32: bits 0..31 -> lo, bit 32 -> hi
// (zero beyond the sub-volume)
__device__ __forceinline__ void mc_inside_row(const mc_params &p, int seg, int y, int z, uint32_t &lo, uint32_t &hi) {
  lo = 0u; hi = 0u;
  if (y >= p.sy || z >= p.sz) return;
  const int nsub = min(33, p.sx - seg * 32);
  const int gx0 = p.lo0 + seg * 32;
  int gy = p.lo1 + y, gz = p.lo2 + z;
  if (gy >= p.c.ny) { if (p.closed) return; gy -= p.c.ny; gz++; }  // the wrapped extra row
  if (gz >= p.c.nz) {                         // the padded extra plane
    if (p.pad_inside) { lo = nsub >= 32 ? 0xffffffffu : ((1u << nsub) - 1u); hi = nsub >= 33; }
    return;
  }
  const int nin = min(nsub, p.c.nx - gx0);    // span bits that lie inside this volume row
  if (nin > 0) {
    const uint32_t *r = p.ibits + ((size_t)gz * p.c.ny + gy) * p.c.w;
    const int wi = gx0 >> 5, sh = gx0 & 31;
    const uint32_t w0 = __ldg(r + wi);
    const uint32_t w1 = (wi + 1 < p.c.w) ? __ldg(r + wi + 1) : 0u;
    lo = __funnelshift_r(w0, w1, sh);
    hi = (w1 >> sh) & 1u;
    if (nin < 33) { hi = 0u; if (nin < 32) lo &= (1u << nin) - 1u; }
  }
  for (int b = nin > 0 ? nin : 0; b < nsub; b++) {  // at most the one wrapped extra column
    const uint32_t in = mc_inside_bit_wrapped(p, gx0 + b, gy, gz);
    if (b < 32) lo |= in << b; else hi = in;
  }
}

__device__ __forceinline__ void mc_flush(const mc_params &p, uint4 *mybuf, unsigned cnt, unsigned lane) {
  __syncwarp();
  unsigned base = 0;
  if (lane == 0) {
    base = atomicAdd(&p.sc->n_active, cnt);
    if (base + cnt > p.active_cap) atomicOr(&p.sc->overflow, 1u);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  for (unsigned i = lane; i < cnt; i += 32)
    if (base + i < p.active_cap) p.active[base + i] = mybuf[i];
  __syncwarp();
}

#ifndef MCB_MINB
#define MCB_MINB 9  /* 56 registers, no spill; measured flat from 8 to 10, slower at 1 (117 registers) and at 12 (spills) */
#endif
__global__ void __launch_bounds__(MCB_THREADS, MCB_MINB) k_mc_classify(const __grid_constant__ mc_params p) {
  __shared__ uint4 rbuf[MCB_THREADS / 32][MCB_BUF];
  const unsigned lane = threadIdx.x & 31;
  uint4 *mybuf = rbuf[threadIdx.x >> 5];
  const long long ncol = (long long)p.sy * p.segs;
  const long long col = (long long)blockIdx.x * MCB_THREADS + threadIdx.x;  // (row y, segment) column
  const bool live = col < ncol;
  const int y = live ? (int)(col / p.segs) : 0;
  const int seg = live ? (int)(col - (long long)y * p.segs) : 0;
  const int z0 = p.zs0 + blockIdx.y * MCB_ZC;
  const int z1 = min(z0 + MCB_ZC, p.zs0 + p.zn);
  const int xb0 = seg * 32;
  const int nvx = live ? min(32, p.sx - xb0) : 0;        // voxels of this segment
  const int nvx1 = live ? min(32, p.sx - 1 - xb0) : 0;   // ... whose x+1 neighbour exists
  const uint32_t vxm = nvx >= 32 ? 0xffffffffu : ((1u << nvx) - 1u);
  const uint32_t vx1m = nvx1 >= 32 ? 0xffffffffu : (nvx1 > 0 ? ((1u << nvx1) - 1u) : 0u);
  const bool vy1 = y + 1 < p.sy;
  uint32_t a = 0, ah = 0, b = 0, bh = 0;  // plane z: row y, row y+1 (h = bit 32 of the span)
  if (live) {
    mc_inside_row(p, seg, y, z0, a, ah);
    mc_inside_row(p, seg, y + 1, z0, b, bh);
  }
  unsigned cnt = 0;  // records parked in mybuf (warp-uniform)
  unsigned long long first = ~0ull;
  for (int z = z0; z < z1; z++) {
    uint32_t c = 0, ch = 0, d = 0, dh = 0;  // plane z+1
    if (live) {
      mc_inside_row(p, seg, y, z + 1, c, ch);
      mc_inside_row(p, seg, y + 1, z + 1, d, dh);
    }
    const bool vz1 = z + 1 < p.sz;
    const uint32_t ax = (a >> 1) | (ah << 31), bx = (b >> 1) | (bh << 31);
    const uint32_t cx = (c >> 1) | (ch << 31), dx = (d >> 1) | (dh << 31);
    const uint32_t exm = (a ^ ax) & vx1m;
    const uint32_t eym = vy1 ? ((a ^ b) & vxm) : 0u;
    const uint32_t ezm = vz1 ? ((a ^ c) & vxm) : 0u;
    const uint32_t orm = a | ax | b | bx | c | cx | d | dx, andm = a & ax & b & bx & c & cx & d & dx;
    const uint32_t trim = (vy1 && vz1) ? ((orm & ~andm) & vx1m) : 0u;
    const uint32_t actm = exm | eym | ezm | trim;
    // where this lane's records go: warp-wide exclusive prefix of the record counts
    const unsigned n = __popc(actm);
    unsigned incl = n;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, dd);
      if (lane >= (unsigned)dd) incl += t;
    }
    const unsigned tot = __shfl_sync(0xffffffffu, incl, 31);
    const size_t row = (size_t)(z - p.zs0) * p.sy + y;  // local row (segment arrays of this rank)
    uint32_t packed = 0;
    if (tot) {
      bool direct = false;
      unsigned gbase = 0;
      if (cnt + tot > MCB_BUF) {
        if (cnt) mc_flush(p, mybuf, cnt, lane);
        cnt = 0;
        if (tot > MCB_BUF) {  // a plane with more records than the buffer holds: straight to the global list
          direct = true;
          if (lane == 0) {
            gbase = atomicAdd(&p.sc->n_active, tot);
            if (gbase + tot > p.active_cap) atomicOr(&p.sc->overflow, 1u);
          }
          gbase = __shfl_sync(0xffffffffu, gbase, 0);
        }
      }
      unsigned slot = (direct ? gbase : cnt) + (incl - n);
      int pv = 0, pt = 0, pc = 0;
      for (uint32_t rest = actm; rest; rest &= rest - 1) {
        const int bb = __ffs(rest) - 1;
        const uint32_t ex = (exm >> bb) & 1u, ey = (eym >> bb) & 1u, ez = (ezm >> bb) & 1u;
        int ntri = 0, hasc = 0, off = 0, lut = 0;
        const int x = xb0 + bb;
        if ((trim >> bb) & 1u) {
          lut = (int)(((a >> bb) & 1u) | (((ax >> bb) & 1u) << 1) | (((bx >> bb) & 1u) << 2) | (((b >> bb) & 1u) << 3) |
                      (((c >> bb) & 1u) << 4) | (((cx >> bb) & 1u) << 5) | (((dx >> bb) & 1u) << 6) | (((d >> bb) & 1u) << 7));
          const uint32_t info = __ldg(p.lutinfo + lut);
          if (info >> 31) {  // ambiguous MC33 case: the face / interior tests need the corner values
            float cv[8];
            cv[0] = mc_data(p, x, y, z); cv[1] = mc_data(p, x + 1, y, z); cv[2] = mc_data(p, x + 1, y + 1, z);
            cv[3] = mc_data(p, x, y + 1, z); cv[4] = mc_data(p, x, y, z + 1); cv[5] = mc_data(p, x + 1, y, z + 1);
            cv[6] = mc_data(p, x + 1, y + 1, z + 1); cv[7] = mc_data(p, x, y + 1, z + 1);
            mc33_select(p.tab, cv, lut, 0, off, ntri, hasc);
          } else {
            off = (int)(info & 0xffffu);
            ntri = (int)(info >> 16);
          }
          if (p.classic && ntri > 0) {
            const unsigned long long key = ((unsigned long long)((size_t)z * p.sy + y) << 16) | (unsigned long long)x;
            first = key < first ? key : first;
          }
        }
        uint4 rec;
        rec.x = (uint32_t)row;
        rec.y = (uint32_t)x | (ex << 16) | (ey << 17) | (ez << 18) | ((uint32_t)ntri << 19) | ((uint32_t)hasc << 23) |
                ((uint32_t)pc << 24);
        rec.z = (uint32_t)pv | ((uint32_t)pt << 7) | ((uint32_t)off << 16);
        rec.w = (uint32_t)lut;
        if (direct) { if (slot < p.active_cap) p.active[slot] = rec; }
        else mybuf[slot] = rec;
        slot++;
        pv += (int)(ex + ey + ez); pt += ntri; pc += hasc;
      }
      packed = (uint32_t)pv | ((uint32_t)pt << 7) | ((uint32_t)pc << 16);
      if (!direct) cnt += tot;
    }
    if (live) {
      p.segbits[row * p.segs + seg] = make_uint4(exm, eym, ezm, packed);
      p.segcnt[row * p.segs + seg] = packed;
    }
    a = c; ah = ch; b = d; bh = dh;
  }
  if (cnt) mc_flush(p, mybuf, cnt, lane);
  if (p.classic) {  // first active cube in raster order: its first soup vertex is the weld's pts[0]
#pragma unroll
    for (int dd = 16; dd; dd >>= 1) {
      unsigned long long o = __shfl_xor_sync(0xffffffffu, first, dd);
      first = o < first ? o : first;
    }
    if (lane == 0 && first != ~0ull) atomicMin(&p.sc->first_cube, first);
  }
}

// scan3: the packed per-segment counts (segbits[].w = nv | nt<<7 | nc<<16) -> exclusive bases
// segbits[].w = vbase, segt[] = tbase, segc[] = cbase; totals -> sc->tot_v/tot_t/tot_c.
#define S3_THREADS 256
#ifndef S3_ITEMS
#define S3_ITEMS 8
#endif
#define S3_TILE (S3_THREADS * S3_ITEMS)
struct u3 { uint32_t v, t, c; };
__device__ __forceinline__ u3 u3_add(u3 a, u3 b) { return {a.v + b.v, a.t + b.t, a.c + b.c}; }
__device__ __forceinline__ u3 u3_unpack(uint32_t w) { return {w & 127u, (w >> 7) & 511u, w >> 16}; }
__device__ __forceinline__ u3 warp_incl_scan3(u3 a) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, a.v, d), t = __shfl_up_sync(0xffffffffu, a.t, d), c = __shfl_up_sync(0xffffffffu, a.c, d);
    if (lane >= (unsigned)d) { a.v += v; a.t += t; a.c += c; }
  }
  return a;
}
// exclusive block scan of one u3 per thread; *total = block sum.  sm: 33 entries
__device__ __forceinline__ u3 block_excl_scan3(u3 a, u3 *total, u3 *sm) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  u3 inc = warp_incl_scan3(a);
  if (lane == 31) sm[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    u3 w = lane < nwarp ? sm[lane] : u3{0, 0, 0};
    u3 winc = warp_incl_scan3(w);
    sm[lane] = {winc.v - w.v, winc.t - w.t, winc.c - w.c};
    if (lane == 31) sm[32] = winc;
  }
  __syncthreads();
  u3 base = sm[warp];
  u3 res = {inc.v - a.v + base.v, inc.t - a.t + base.t, inc.c - a.c + base.c};
  *total = sm[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(S3_THREADS) k_scan3_reduce(const uint32_t *__restrict__ segcnt, size_t n, uint32_t *__restrict__ part,
                                                             size_t nblk) {
  __shared__ u3 sm[33];
  size_t base = (size_t)blockIdx.x * S3_TILE;
  u3 s = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < S3_ITEMS; i++) {
    size_t idx = base + (size_t)i * S3_THREADS + threadIdx.x;
    if (idx < n) s = u3_add(s, u3_unpack(__ldg(segcnt + idx)));
  }
  u3 tot;
  block_excl_scan3(s, &tot, sm);
  if (threadIdx.x == 0) { part[blockIdx.x] = tot.v; part[nblk + blockIdx.x] = tot.t; part[2 * nblk + blockIdx.x] = tot.c; }
}

// one block: exclusive scan of the three partial arrays in place; totals -> tot[0..2]
__global__ void __launch_bounds__(1024) k_scan3_single(uint32_t *part, size_t nblk, unsigned int *tot) {
  __shared__ u3 sm[33];
  u3 carry = {0, 0, 0};
  for (size_t base = 0; base < nblk; base += blockDim.x) {
    size_t idx = base + threadIdx.x;
    u3 v = idx < nblk ? u3{part[idx], part[nblk + idx], part[2 * nblk + idx]} : u3{0, 0, 0};
    u3 t;
    u3 ex = block_excl_scan3(v, &t, sm);
    if (idx < nblk) { part[idx] = ex.v + carry.v; part[nblk + idx] = ex.t + carry.t; part[2 * nblk + idx] = ex.c + carry.c; }
    carry = u3_add(carry, t);
  }
  if (threadIdx.x == 0) { tot[0] = carry.v; tot[1] = carry.t; tot[2] = carry.c; }
}

// A base is only ever read for a segment whose own count is non-zero (a vertex / triangle / centroid of that segment is
// being numbered), so only those are written: ~5 % of the segments of a smooth volume instead of three scattered
// 4-byte stores per segment.  force_idx: one more vertex base that is read directly (first segment of the second own
// plane = slabs' n_first).
/* measured on G1024: 0.326 ms at 1 (all the registers it wants), 0.354 unconstrained, 0.368 at 8 */
#ifndef S3A_MINB
#define S3A_MINB 1
#endif
__global__ void __launch_bounds__(S3_THREADS, S3A_MINB) k_scan3_apply(uint4 *__restrict__ seg, const uint32_t *__restrict__ segcnt, size_t n,
                                                            const uint32_t *__restrict__ part, size_t nblk,
                                                            uint32_t *__restrict__ segt, uint32_t *__restrict__ segc, uint32_t voff,
                                                            size_t force_idx) {
  __shared__ u3 sm[33];
  size_t base = (size_t)blockIdx.x * S3_TILE + (size_t)threadIdx.x * S3_ITEMS;
  uint32_t raw[S3_ITEMS];
  if (base + S3_ITEMS <= n && (reinterpret_cast<uintptr_t>(segcnt + base) & 15) == 0) {  // 8 consecutive counts = one 32-byte sector per thread
#pragma unroll
    for (int q = 0; q < S3_ITEMS / 4; q++) {
      const uint4 a = __ldg(reinterpret_cast<const uint4 *>(segcnt + base) + q);
      raw[4 * q] = a.x; raw[4 * q + 1] = a.y; raw[4 * q + 2] = a.z; raw[4 * q + 3] = a.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < S3_ITEMS; i++) raw[i] = base + i < n ? __ldg(segcnt + base + i) : 0u;
  }
  u3 v[S3_ITEMS];
  u3 s = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < S3_ITEMS; i++) {
    v[i] = u3_unpack(raw[i]);
    s = u3_add(s, v[i]);
  }
  u3 tot;
  u3 ex = block_excl_scan3(s, &tot, sm);
  ex = u3_add(ex, u3{part[blockIdx.x] + voff, part[nblk + blockIdx.x], part[2 * nblk + blockIdx.x]});
#pragma unroll
  for (int i = 0; i < S3_ITEMS; i++) {
    size_t idx = base + i;
    if (idx < n) {
      if (v[i].v || idx == force_idx) seg[idx].w = ex.v;
      if (v[i].t) segt[idx] = ex.t;
      if (v[i].c) segc[idx] = ex.c;
    }
    ex = u3_add(ex, v[i]);
  }
}

// totals first (the vertex base of a rank is the sum of the lower ranks' totals), bases second
static int mc_scan3_totals(b2m_ctx *ctx, mc_params &p, size_t nseg, b2m_scalars *d_sc) {
  if (nseg == 0) { CU_TRY(cudaMemsetAsync(&d_sc->tot_v, 0, 12, ctx->stream)); return B2M_OK; }
  size_t nblk = (nseg + S3_TILE - 1) / S3_TILE;
  B2M_TRY(b2m_reserve(ctx, BUF_SCAN1, nblk * 3 * 4));
  uint32_t *part = b2m_ptr<uint32_t>(ctx, BUF_SCAN1);
  KT_LAUNCH(ctx, "scan3_reduce", k_scan3_reduce<<<(unsigned)nblk, S3_THREADS, 0, ctx->stream>>>(p.segcnt, nseg, part, nblk));
  KT_LAUNCH(ctx, "scan3_single", k_scan3_single<<<1, 1024, 0, ctx->stream>>>(part, nblk, &d_sc->tot_v));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}
static int mc_scan3_apply(b2m_ctx *ctx, mc_params &p, size_t nseg, uint32_t voff, size_t force_idx) {
  if (nseg == 0) return B2M_OK;
  size_t nblk = (nseg + S3_TILE - 1) / S3_TILE;
  uint32_t *part = b2m_ptr<uint32_t>(ctx, BUF_SCAN1);
  KT_LAUNCH(ctx, "scan3_apply", k_scan3_apply<<<(unsigned)nblk, S3_THREADS, 0, ctx->stream>>>(p.segbits, p.segcnt, nseg, part, nblk, p.segt, p.segc, voff, force_idx));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

// ---- pass B: emit ------------------------------------------------------------------------------
struct mc_emit_params {
  double *verts;         // own vertices: [edge block | centroid block]
  int *tris;             // own triangles (global vertex ids)
  b2m_item *items;       // weld items (vertices / soup copies near a grid corner)
  uint32_t *nearbits;    // bit per own edge vertex: within B2M_NEAR_TOL of a grid corner (zeroed before the launch)
  unsigned int item_cap;
  unsigned int n_active;
  unsigned int e_off;    // global id of the first own edge vertex (segbits[].w already includes it)
  unsigned int nv_edge_l;  // own edge vertices (the own centroid block starts there in verts[])
  unsigned int c_base;   // global id of the first own centroid vertex = global edge-vertex total + lower ranks' centroids
  unsigned int t_off;    // global index of the first own triangle
  unsigned long long first_cube;  // classic: (global row << 16 | x) of the first active cube of the volume
};

// index of the edge vertex owned by sub-volume voxel (x,row) on `axis`
__device__ __forceinline__ uint32_t mc_vidx(const mc_params &p, size_t row, int x, int axis) {
  size_t s = row * p.segs + (x >> 5);
  uint4 b = __ldg(p.segbits + s);
  unsigned lane = x & 31, m = (1u << lane) - 1u;
  uint32_t n = b.w + __popc(b.x & m) + __popc(b.y & m) + __popc(b.z & m);
  if (axis >= 1) n += (b.x >> lane) & 1u;
  if (axis == 2) n += (b.y >> lane) & 1u;
  return n;
}

__device__ __forceinline__ void push_item(const mc_params &p, const mc_emit_params &e, double x, double y, double z, uint32_t id,
                                          uint32_t vid) {
  unsigned pos = atomicAdd(&p.sc->n_cand, 1u);
  if (pos < e.item_cap) {
    b2m_item it;
    it.pos[0] = x; it.pos[1] = y; it.pos[2] = z; it.id = id; it.vid = vid;
    e.items[pos] = it;
  } else {
    atomicOr(&p.sc->overflow, 16u);
  }
}

// Lewiner edge vertex: u = c0/(c0-c1) (0.5 when the denominator is zero), local f32 coordinate
// (float)i + u, exported as (double)(float)(local + (float)lo)   (src/MarchingCubes.c:943-949,:1127-1129)
__device__ __forceinline__ float lew_u(float c0, float c1) {
  float den = __fsub_rn(c0, c1);
  return den != 0.0f ? __fdiv_rn(c0, den) : 0.5f;
}
__device__ __forceinline__ bool near_int(float f, float tol) { return fabsf(__fsub_rn(f, rintf(f))) < tol; }
__device__ __forceinline__ void flag_near(const mc_emit_params &e, uint32_t local) { atomicOr(e.nearbits + (local >> 5), 1u << (local & 31u)); }
__device__ __forceinline__ bool near_int_d(double f, double tol) { return fabs(f - rint(f)) < tol; }

/* CTAs per SM the register budget is cut for.  The kernel waits on dependent L2 round trips, so resident warps beat
   spill-free code: Lewiner 3.26 ms of marching cubes at 16 (32 registers, spills), 3.31 at 12, 3.44 at 10 (48 registers,
   no spills); classic keeps 12 (FP64 positions) */
#ifndef EMIT_MINB
#define EMIT_MINB (CLASSIC ? 12 : 16)
#endif
// EMIT_STAGE=1 (compiled out: measured slower): Lewiner output through a per-warp shared-memory stage.  Consecutive
// active records are consecutive voxels of a row, so the vertices / triangles of a warp's 32 records are a few
// contiguous runs of the output arrays; written straight from the threads that compute them they are 8- and 4-byte
// stores at 24..72-byte strides.  Staged, a store instruction covers 256 (vertices) / 128 (triangles) contiguous bytes:
// 335 M -> 79 M L1 store sectors on G1024 - and 2.07 -> 2.49 ms, with the same DRAM traffic.  The write sectors that
// crowded L2 were not these stores but the local-memory array of edge ids (now shared memory, see ev() below).
#ifndef EMIT_STAGE
#define EMIT_STAGE 0
#endif
#define EM_VCAP 96   /* 32 records x 3 edge vertices */
#define EM_TCAP 128  /* triangles staged per warp; a warp with more (of 32 x 12 possible) stores them directly */
// CLASSIC as a template parameter: with a run-time flag the Lewiner path carried the classic path's dynamically indexed
// corner array (local memory) and its FP64 registers
template <bool CLASSIC>
__global__ void __launch_bounds__(128, EMIT_MINB) k_mc_emit(mc_params p, mc_emit_params e) {
#if EMIT_STAGE
  __shared__ float s_v[4][EM_VCAP * 3];
  __shared__ uint32_t s_vg[4][EM_VCAP];
  __shared__ int s_t[4][EM_TCAP * 3];
  __shared__ uint32_t s_tg[4][EM_TCAP];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#endif
  __shared__ uint32_t s_ev[13][128];
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool staged = EMIT_STAGE && !CLASSIC;  // block-uniform
  const bool live = i < e.n_active;
  if (!live && !staged) return;
  const uint4 r = live ? p.active[i] : make_uint4(0u, 0u, 0u, 0u);  // a dead lane of a staged warp: no edges, no triangles
  const size_t row = r.x;
  const int x = r.y & 0xffff;
  const bool ex = (r.y >> 16) & 1, ey = (r.y >> 17) & 1, ez = (r.y >> 18) & 1;
  const int ntri = (r.y >> 19) & 15, hasc = (r.y >> 23) & 1, pc = (r.y >> 24) & 31;
  const int pv = r.z & 127, pt = (r.z >> 7) & 511, off = (int)(r.z >> 16);
  const int zl = (int)(row / p.sy), y = (int)(row - (size_t)zl * p.sy);
  const int z = zl + p.zs0;  // sub-volume plane (global); row / sidx index this rank's segment arrays
  const size_t sidx = row * p.segs + (x >> 5);
  // corner values (only the ones that exist inside the sub-volume)
  const bool vx1 = live && x + 1 < p.sx, vy1 = live && y + 1 < p.sy, vz1 = live && z + 1 < p.sz;
  float c[8];
  c[0] = live ? mc_data(p, x, y, z) : 0.f;
  c[1] = vx1 ? mc_data(p, x + 1, y, z) : c[0];
  c[3] = vy1 ? mc_data(p, x, y + 1, z) : c[0];
  c[4] = vz1 ? mc_data(p, x, y, z + 1) : c[0];
  if (ntri && (CLASSIC || hasc)) {  // Lewiner needs the far corners only for a centroid vertex
    c[2] = mc_data(p, x + 1, y + 1, z);
    c[5] = mc_data(p, x + 1, y, z + 1);
    c[6] = mc_data(p, x + 1, y + 1, z + 1);
    c[7] = mc_data(p, x, y + 1, z + 1);
  } else {
    c[2] = c[5] = c[6] = c[7] = c[0];
  }
  const float flo0 = (float)(p.lo0 + p.org0), flo1 = (float)(p.lo1 + p.org1), flo2 = (float)(p.lo2 + p.org2);
  // ---- vertex ids of the 12 cube edges (src/MarchingCubes.c:813-825) ----
  // Looked up BEFORE the vertex stores below (r1: validated on the single-volume and slab suites, 2.10 -> 2.06 ms): behind them (and behind the `ntri == 0` exit) the segment-record loads
  // started only after the corner values had arrived and the vertices were written - a third dependent round trip
  // through L2 per thread.  One record load serves every edge that shares a (row, segment): <= 4 loads instead of 12.
  // the 12 edge-vertex ids (+ the centroid's) are indexed by table entries: as a per-thread array they lived in LOCAL
  // memory - 13 write-through stores and up to 36 loads per thread, 70 % of the kernel's L2 write sectors (ncu r2:
  // 289 M write sectors with the output itself staged down to 79 M) and 1.3 GB of DRAM write-backs.  Shared memory,
  // one column per thread (bank = lane).
#define ev(k) s_ev[(k)][threadIdx.x]
  const unsigned lut = r.w;  // inside bits of the corners = the cube index the classify pass built from the inside-bit rows
  const bool in0 = lut & 1u, in1 = (lut >> 1) & 1u, in2 = (lut >> 2) & 1u, in3 = (lut >> 3) & 1u;
  const bool in4 = (lut >> 4) & 1u, in5 = (lut >> 5) & 1u, in6 = (lut >> 6) & 1u, in7 = (lut >> 7) & 1u;
  if (ntri) {
    const bool c0 = in0 != in1, c1 = in1 != in2, c2 = in3 != in2, c3 = in0 != in3, c4 = in4 != in5, c5 = in5 != in6;
    const bool c6 = in7 != in6, c7 = in4 != in7, c8 = in0 != in4, c9 = in1 != in5, c10 = in2 != in6, c11 = in3 != in7;
    if (((x + 1) & 31) != 0) {  // x and x+1 share their segment (31 voxels of 32): four record loads at most
      const uint4 none = make_uint4(0u, 0u, 0u, 0u);
      const uint4 *sA = p.segbits + row * p.segs + (x >> 5), *sB = sA + p.segs, *sC = sA + (size_t)p.sy * p.segs, *sD = sC + p.segs;
      const uint4 A = (c0 | c3 | c8 | c1 | c9) ? __ldg(sA) : none;
      const uint4 B = (c2 | c11 | c10) ? __ldg(sB) : none;
      const uint4 Cr = (c4 | c7 | c5) ? __ldg(sC) : none;
      const uint4 D = c6 ? __ldg(sD) : none;
      // mc_vidx on a record already loaded: vertices before voxel xx in the segment, + the x (+ y) vertex of xx itself
      const unsigned l0 = x & 31, m0 = (1u << l0) - 1u, l1 = l0 + 1, m1 = (1u << l1) - 1u;
      const uint32_t a0 = A.w + __popc(A.x & m0) + __popc(A.y & m0) + __popc(A.z & m0);
      const uint32_t a1 = A.w + __popc(A.x & m1) + __popc(A.y & m1) + __popc(A.z & m1);
      const uint32_t b0 = B.w + __popc(B.x & m0) + __popc(B.y & m0) + __popc(B.z & m0);
      const uint32_t b1 = B.w + __popc(B.x & m1) + __popc(B.y & m1) + __popc(B.z & m1);
      const uint32_t q0 = Cr.w + __popc(Cr.x & m0) + __popc(Cr.y & m0) + __popc(Cr.z & m0);
      const uint32_t q1 = Cr.w + __popc(Cr.x & m1) + __popc(Cr.y & m1) + __popc(Cr.z & m1);
      const uint32_t d0 = D.w + __popc(D.x & m0) + __popc(D.y & m0) + __popc(D.z & m0);
      ev(0) = c0 ? a0 : 0xffffffffu;
      ev(1) = c1 ? a1 + ((A.x >> l1) & 1u) : 0xffffffffu;
      ev(2) = c2 ? b0 : 0xffffffffu;
      ev(3) = c3 ? a0 + ((A.x >> l0) & 1u) : 0xffffffffu;
      ev(4) = c4 ? q0 : 0xffffffffu;
      ev(5) = c5 ? q1 + ((Cr.x >> l1) & 1u) : 0xffffffffu;
      ev(6) = c6 ? d0 : 0xffffffffu;
      ev(7) = c7 ? q0 + ((Cr.x >> l0) & 1u) : 0xffffffffu;
      ev(8) = c8 ? a0 + ((A.x >> l0) & 1u) + ((A.y >> l0) & 1u) : 0xffffffffu;
      ev(9) = c9 ? a1 + ((A.x >> l1) & 1u) + ((A.y >> l1) & 1u) : 0xffffffffu;
      ev(10) = c10 ? b1 + ((B.x >> l1) & 1u) + ((B.y >> l1) & 1u) : 0xffffffffu;
      ev(11) = c11 ? b0 + ((B.x >> l0) & 1u) + ((B.y >> l0) & 1u) : 0xffffffffu;
    } else {  // x+1 opens the next segment of its row: one lookup per edge
      const size_t rowY = row + 1, rowZ = row + p.sy, rowYZ = row + p.sy + 1;
      ev(0) = c0 ? mc_vidx(p, row, x, 0) : 0xffffffffu;
      ev(1) = c1 ? mc_vidx(p, row, x + 1, 1) : 0xffffffffu;
      ev(2) = c2 ? mc_vidx(p, rowY, x, 0) : 0xffffffffu;
      ev(3) = c3 ? mc_vidx(p, row, x, 1) : 0xffffffffu;
      ev(4) = c4 ? mc_vidx(p, rowZ, x, 0) : 0xffffffffu;
      ev(5) = c5 ? mc_vidx(p, rowZ, x + 1, 1) : 0xffffffffu;
      ev(6) = c6 ? mc_vidx(p, rowYZ, x, 0) : 0xffffffffu;
      ev(7) = c7 ? mc_vidx(p, rowZ, x, 1) : 0xffffffffu;
      ev(8) = c8 ? mc_vidx(p, row, x, 2) : 0xffffffffu;
      ev(9) = c9 ? mc_vidx(p, row, x + 1, 2) : 0xffffffffu;
      ev(10) = c10 ? mc_vidx(p, rowY, x + 1, 2) : 0xffffffffu;
      ev(11) = c11 ? mc_vidx(p, rowY, x, 2) : 0xffffffffu;
    }
    ev(12) = 0xffffffffu;
  }
  // ---- own edge vertices ----
#if EMIT_STAGE
  // exclusive offsets of this lane's edge vertices / triangles in the warp's stage (one packed scan: sums <= 96 / 480)
  unsigned voff = 0, toff = 0, vtot = 0, ttot = 0;
  if (staged) {
    const unsigned mine = (unsigned)(ex + ey + ez) | ((unsigned)ntri << 16);
    unsigned inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned up = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (unsigned)d) inc += up;
    }
    const unsigned tot = __shfl_sync(0xffffffffu, inc, 31);
    voff = (inc - mine) & 0xffffu; toff = (inc - mine) >> 16;
    vtot = tot & 0xffffu; ttot = tot >> 16;
  }
#endif
  if (ex | ey | ez) {
    uint32_t vid = __ldg(&p.segbits[sidx].w) + (uint32_t)pv;
    if (!CLASSIC) {
      const float fx = (float)x, fy = (float)y, fz = (float)z;
      const float ox = __fadd_rn(fx, flo0), oy = __fadd_rn(fy, flo1), oz = __fadd_rn(fz, flo2);
      // one edge vertex: into the warp's stage (or straight to verts[]), flag / weld item when it sits near a grid corner
      auto put = [&](float qx, float qy, float qz, float qfree) {
#if EMIT_STAGE
        float *sv = s_v[wid] + 3 * voff;
        sv[0] = qx; sv[1] = qy; sv[2] = qz;
        s_vg[wid][voff] = vid - e.e_off;
        voff++;
#else
        double *o = e.verts + 3 * (size_t)(vid - e.e_off);
        o[0] = (double)qx; o[1] = (double)qy; o[2] = (double)qz;
#endif
        if (near_int(qfree, B2M_NEAR_TOL)) {
          flag_near(e, vid - e.e_off);
          if (near_int(qfree, 2e-5f)) push_item(p, e, (double)qx, (double)qy, (double)qz, vid, vid);
        }
        vid++;
      };
      if (ex) { const float px = __fadd_rn(__fadd_rn(fx, lew_u(c[0], c[1])), flo0); put(px, oy, oz, px); }
      if (ey) { const float py = __fadd_rn(__fadd_rn(fy, lew_u(c[0], c[3])), flo1); put(ox, py, oz, py); }
      if (ez) { const float pz = __fadd_rn(__fadd_rn(fz, lew_u(c[0], c[4])), flo2); put(ox, oy, pz, pz); }
    } else {
      // classic: FP64, mu = (iso - v1)/(v2 - v1), p = p1 + mu*(p2-p1) (src/oldcubes.c:35-38) with the
      // direction of the LAST cube (raster order) that touches the edge: that soup copy has the
      // highest index, so its coordinates survive the reference's weld (src/meshify.c:99-100).
      const double iso = (double)p.c.iso;
      const double gx = (double)(p.lo0 + p.org0 + x), gy = (double)(p.lo1 + p.org1 + y), gz = (double)(p.lo2 + p.org2 + z);
      const double v0 = (double)c[0];
      if (ex) {
        const double v1 = (double)c[1];
        // touching cubes, last in raster order: layer (z or z-1), then row y if valid (edge 0/4: +x) else y-1 (edge 2/6: -x)
        const bool plus = y <= p.sy - 2;
        double px = plus ? __dadd_rn(gx, __ddiv_rn(__dsub_rn(iso, v0), __dsub_rn(v1, v0)))
                         : __dadd_rn(gx + 1.0, __dmul_rn(__ddiv_rn(__dsub_rn(iso, v1), __dsub_rn(v0, v1)), -1.0));
        double *o = e.verts + 3 * (size_t)(vid - e.e_off);
        o[0] = px; o[1] = gy; o[2] = gz;
        if (near_int_d(px, (double)B2M_NEAR_TOL)) flag_near(e, vid - e.e_off);
        vid++;
      }
      if (ey) {
        const double v1 = (double)c[3];
        // last toucher: column x if valid (edge 3/7: from y+1 down to y, "-y") else x-1 (edge 1/5: +y)
        const bool minus = x <= p.sx - 2;
        double py = minus ? __dadd_rn(gy + 1.0, __dmul_rn(__ddiv_rn(__dsub_rn(iso, v1), __dsub_rn(v0, v1)), -1.0))
                          : __dadd_rn(gy, __ddiv_rn(__dsub_rn(iso, v0), __dsub_rn(v1, v0)));
        double *o = e.verts + 3 * (size_t)(vid - e.e_off);
        o[0] = gx; o[1] = py; o[2] = gz;
        if (near_int_d(py, (double)B2M_NEAR_TOL)) flag_near(e, vid - e.e_off);
        vid++;
      }
      if (ez) {
        const double v1 = (double)c[4];
        double pz = __dadd_rn(gz, __ddiv_rn(__dsub_rn(iso, v0), __dsub_rn(v1, v0)));
        double *o = e.verts + 3 * (size_t)(vid - e.e_off);
        o[0] = gx; o[1] = gy; o[2] = pz;
        if (near_int_d(pz, (double)B2M_NEAR_TOL)) flag_near(e, vid - e.e_off);
      }
    }
  }
#if EMIT_STAGE
  if (staged) {  // the warp's edge vertices, 32 consecutive doubles (256 bytes inside a run) per store instruction
    __syncwarp();
    for (unsigned wd = lane; wd < 3u * vtot; wd += 32) {
      const unsigned j = (wd * 43691u) >> 17;  // wd / 3 (wd < 288)
      e.verts[3 * (size_t)s_vg[wid][j] + (wd - 3u * j)] = (double)s_v[wid][wd];
    }
    if (ttot == 0) return;  // warp-uniform
  } else
#endif
  if (!ntri) return;
  if (ntri && hasc) {
    // centroid of the cube's existing edge vertices, summed in edge-code order in f32 local
    // coordinates, divided by the f32 count (src/MarchingCubes.c:1042-1071); then + lo in f32.
    const float fx = (float)x, fy = (float)y, fz = (float)z;
    const float fx1 = (float)(x + 1), fy1 = (float)(y + 1), fz1 = (float)(z + 1);
    float sx = 0.f, sy = 0.f, sz = 0.f, u = 0.f;
#define ACC(px_, py_, pz_) do { u = __fadd_rn(u, 1.0f); sx = __fadd_rn(sx, (px_)); sy = __fadd_rn(sy, (py_)); sz = __fadd_rn(sz, (pz_)); } while (0)
    if (in0 != in1) ACC(__fadd_rn(fx, lew_u(c[0], c[1])), fy, fz);
    if (in1 != in2) ACC(fx1, __fadd_rn(fy, lew_u(c[1], c[2])), fz);
    if (in3 != in2) ACC(__fadd_rn(fx, lew_u(c[3], c[2])), fy1, fz);
    if (in0 != in3) ACC(fx, __fadd_rn(fy, lew_u(c[0], c[3])), fz);
    if (in4 != in5) ACC(__fadd_rn(fx, lew_u(c[4], c[5])), fy, fz1);
    if (in5 != in6) ACC(fx1, __fadd_rn(fy, lew_u(c[5], c[6])), fz1);
    if (in7 != in6) ACC(__fadd_rn(fx, lew_u(c[7], c[6])), fy1, fz1);
    if (in4 != in7) ACC(fx, __fadd_rn(fy, lew_u(c[4], c[7])), fz1);
    if (in0 != in4) ACC(fx, fy, __fadd_rn(fz, lew_u(c[0], c[4])));
    if (in1 != in5) ACC(fx1, fy, __fadd_rn(fz, lew_u(c[1], c[5])));
    if (in2 != in6) ACC(fx1, fy1, __fadd_rn(fz, lew_u(c[2], c[6])));
    if (in3 != in7) ACC(fx, fy1, __fadd_rn(fz, lew_u(c[3], c[7])));
#undef ACC
    if (u > 0.f) { sx = __fdiv_rn(sx, u); sy = __fdiv_rn(sy, u); sz = __fdiv_rn(sz, u); }
    const float ox = __fadd_rn(sx, flo0), oy = __fadd_rn(sy, flo1), oz = __fadd_rn(sz, flo2);
    const uint32_t cl = __ldg(p.segc + sidx) + (uint32_t)pc;  // index in the own centroid block
    const uint32_t vid = e.c_base + cl;
    double *o = e.verts + 3 * (size_t)(e.nv_edge_l + cl);
    o[0] = (double)ox; o[1] = (double)oy; o[2] = (double)oz;
    if (near_int(ox, 1e-4f) || near_int(oy, 1e-4f) || near_int(oz, 1e-4f)) push_item(p, e, o[0], o[1], o[2], vid, vid);
    ev(12) = vid;
  }
  if (CLASSIC && (((unsigned long long)((size_t)z * p.sy + y) << 16) | (unsigned long long)x) == e.first_cube) {
    // pts[0] of the reference's soup: first table edge of the first active cube, interpolated in
    // THAT cube's direction (src/oldcubes.c:428-451, :22-40)
    const int a0 = p.tab[off];
    const int ca = MC_EDGE_A[a0], cb = MC_EDGE_B[a0];
    const double iso = (double)p.c.iso;
    const double mu = __ddiv_rn(__dsub_rn(iso, (double)c[ca]), __dsub_rn((double)c[cb], (double)c[ca]));
    const double ax = (double)(p.lo0 + p.org0 + x + ((ca ^ (ca >> 1)) & 1)), ay = (double)(p.lo1 + p.org1 + y + ((ca >> 1) & 1)), az = (double)(p.lo2 + p.org2 + z + (ca >> 2));
    const double bx = (double)(p.lo0 + p.org0 + x + ((cb ^ (cb >> 1)) & 1)), by = (double)(p.lo1 + p.org1 + y + ((cb >> 1) & 1)), bz = (double)(p.lo2 + p.org2 + z + (cb >> 2));
    p.sc->pts0[0] = __dadd_rn(ax, __dmul_rn(mu, bx - ax));
    p.sc->pts0[1] = __dadd_rn(ay, __dmul_rn(mu, by - ay));
    p.sc->pts0[2] = __dadd_rn(az, __dmul_rn(mu, bz - az));
  }
  // ---- triangles ----
  const size_t tl0 = ntri ? (size_t)__ldg(p.segt + sidx) + (size_t)pt : 0;  // index in the own triangle array
  const signed char *tl = p.tab + off;
#if EMIT_STAGE
  if (staged && ttot <= EM_TCAP) {  // warp-uniform
    int *st = s_t[wid] + 3 * toff;
    for (int k = 0; k < ntri; k++) {
      const int a = tl[3 * k], b = tl[3 * k + 1], cc = tl[3 * k + 2];
      st[3 * k] = (int)ev(cc); st[3 * k + 1] = (int)ev(b); st[3 * k + 2] = (int)ev(a);  // reversed (:1134-1136)
      s_tg[wid][toff + k] = (uint32_t)(tl0 + k);
    }
    __syncwarp();
    for (unsigned wd = lane; wd < 3u * ttot; wd += 32) {
      const unsigned j = (wd * 43691u) >> 17;  // wd / 3 (wd < 384)
      e.tris[3 * (size_t)s_tg[wid][j] + (wd - 3u * j)] = s_t[wid][wd];
    }
    return;
  }
#endif
  int *t = e.tris + 3 * tl0;
  for (int k = 0; k < ntri; k++) {
    int a = tl[3 * k], b = tl[3 * k + 1], cc = tl[3 * k + 2];
    if (CLASSIC) { t[3 * k] = (int)ev(a); t[3 * k + 1] = (int)ev(b); t[3 * k + 2] = (int)ev(cc); }
    else { t[3 * k] = (int)ev(cc); t[3 * k + 1] = (int)ev(b); t[3 * k + 2] = (int)ev(a); }  // reversed (:1134-1136)
  }
  if (CLASSIC) {
    // weld items of the classic back-end are SOUP COPIES (the reference welds the soup, so equal keys are
    // ordered by soup index and an edge vertex can be split between clusters, src/meshify.c:60-79):
    // every triangle corner whose position, interpolated in THIS cube's direction (src/oldcubes.c:22-40,
    // :428-451), lies within 2e-5 of a grid corner.  f32 pre-filter per edge, exact FP64 only for the few.
    const double iso = (double)p.c.iso;
    unsigned nearm = 0;
#pragma unroll
    for (int a = 0; a < 12; a++) {
      const float va = c[MC_EDGE_A[a]], vb = c[MC_EDGE_B[a]];
      const float den = fabsf(vb - va) * 1e-4f;
      if (fabsf(p.c.iso - va) <= den || fabsf(p.c.iso - vb) <= den) nearm |= 1u << a;
    }
    if (nearm) {
      for (int k = 0; k < ntri; k++) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
          const int a = tl[3 * k + q];
          if (!((nearm >> a) & 1u)) continue;
          const int ca = MC_EDGE_A[a], cb = MC_EDGE_B[a];
          const double mu = __ddiv_rn(__dsub_rn(iso, (double)c[ca]), __dsub_rn((double)c[cb], (double)c[ca]));
          const double ax = (double)(p.lo0 + p.org0 + x + ((ca ^ (ca >> 1)) & 1)), ay = (double)(p.lo1 + p.org1 + y + ((ca >> 1) & 1)), az = (double)(p.lo2 + p.org2 + z + (ca >> 2));
          const double bx = (double)(p.lo0 + p.org0 + x + ((cb ^ (cb >> 1)) & 1)), by = (double)(p.lo1 + p.org1 + y + ((cb >> 1) & 1)), bz = (double)(p.lo2 + p.org2 + z + (cb >> 2));
          const double px = __dadd_rn(ax, __dmul_rn(mu, bx - ax)), py = __dadd_rn(ay, __dmul_rn(mu, by - ay)), pz = __dadd_rn(az, __dmul_rn(mu, bz - az));
          const double free_axis = a >= 8 ? pz : ((a & 1) ? py : px);
          if (near_int_d(free_axis, 2e-5))
            push_item(p, e, px, py, pz, 3u * (uint32_t)(e.t_off + tl0 + k) + (uint32_t)q, ev(a));
        }
      }
    }
  }
}

#undef ev

// Per-cube-index summary tables derived from the case tables, uploaded after the blob:
//   info = table offset | ntri << 16 | (needs the MC33 face/interior tests) << 31
// [0..255] Lewiner MC33 (src/MarchingCubes.c:479-794, unambiguous cases resolved here), [256..511]
// the classic table (oldcubes.c / Lewiner originalMC, src/MarchingCubes.c:471-477).
#define MCT_INFO_OFF ((MCT_TOTAL + 15) & ~15)
static int upload_tables(b2m_ctx *ctx) {
  if (ctx->tables_ready) return B2M_OK;
  extern const signed char *b2m_mc_table_blob(void);
  const signed char *blob = b2m_mc_table_blob();
  static uint32_t info[512];
  for (int lut = 0; lut < 256; lut++) {
    const int kase = blob[MCT_cases + 2 * lut], cfg = blob[MCT_cases + 2 * lut + 1];
    uint32_t v = 0;
    switch (kase) {
      case 0: v = 0; break;
      case 1: v = (uint32_t)(MCT_tiling1 + cfg * MCT_tiling1_ROW) | (1u << 16); break;
      case 2: v = (uint32_t)(MCT_tiling2 + cfg * MCT_tiling2_ROW) | (2u << 16); break;
      case 5: v = (uint32_t)(MCT_tiling5 + cfg * MCT_tiling5_ROW) | (3u << 16); break;
      case 8: v = (uint32_t)(MCT_tiling8 + cfg * MCT_tiling8_ROW) | (2u << 16); break;
      case 9: v = (uint32_t)(MCT_tiling9 + cfg * MCT_tiling9_ROW) | (4u << 16); break;
      case 11: v = (uint32_t)(MCT_tiling11 + cfg * MCT_tiling11_ROW) | (4u << 16); break;
      case 14: v = (uint32_t)(MCT_tiling14 + cfg * MCT_tiling14_ROW) | (4u << 16); break;
      default: v = 0x80000000u; break;  // 3,4,6,7,10,12,13: decided per cube by mc33_select
    }
    info[lut] = v;
    int o = MCT_casesClassic + 16 * lut, n = 0;
    while (n < 5 && blob[o + 3 * n] != -1) n++;
    info[256 + lut] = (uint32_t)o | ((uint32_t)n << 16);
  }
  B2M_TRY(b2m_reserve(ctx, BUF_TABLES, MCT_INFO_OFF + sizeof(info)));
  char *d = b2m_ptr<char>(ctx, BUF_TABLES);
  CU_TRY(cudaMemcpyAsync(d, blob, MCT_TOTAL, cudaMemcpyHostToDevice, ctx->stream));
  CU_TRY(cudaMemcpyAsync(d + MCT_INFO_OFF, info, sizeof(info), cudaMemcpyHostToDevice, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->tables_ready = 1;
  return B2M_OK;
}

// Slabs: sub-volume planes are owned by the rank that owns the matching volume plane (the last rank also owns
// the Lewiner pad plane); vertex ids are global: rank r's edge vertices follow those of ranks < r, all centroid
// vertices follow all edge vertices (the reference's emission order over the whole volume).
int b2m_mc_run(b2m_ctx *ctx, b2m_comm *comm, const b2m_slab &sl, const b2m_geom &g, const b2m_opts *o,
               const b2m_front_out *fo, b2m_mesh_dev *mesh) {
  B2M_TRY(upload_tables(ctx));
  b2m_scalars *d_sc = b2m_ptr<b2m_scalars>(ctx, BUF_SCALARS);
  const int W = sl.world;
  mc_params p;
  memset(&p, 0, sizeof(p));
  // EXT buffers shifted so that kernels index them with global z
  const long long vshift = (long long)sl.ez0 * g.nxy, wshift = (long long)sl.ez0 * g.ny * g.w;
  p.c.S = fo->S - vshift;
  p.c.fill = fo->fill ? fo->fill - wshift : nullptr;
  p.c.keep = fo->keep ? fo->keep - wshift : nullptr;
  p.ibits = fo->ibits - wshift;
  p.c.nx = g.nx; p.c.ny = g.ny; p.c.nz = sl.gnz; p.c.w = g.w;
  p.c.iso = fo->iso; p.c.mn = fo->vmin; p.c.edge_max = fo->edge_max;
  p.classic = o->backend == B2M_BACKEND_CLASSIC;
  p.original_mc = o->original_mc != 0;
  p.lo0 = fo->lo[0]; p.lo1 = fo->lo[1]; p.lo2 = fo->lo[2];
  p.org0 = ctx->origin[0]; p.org1 = ctx->origin[1]; p.org2 = ctx->origin[2];
  if (p.classic) {  // voxels lo .. hi-1 (src/oldcubes.c:475-478)
    p.sx = fo->hi[0] - fo->lo[0]; p.sy = fo->hi[1] - fo->lo[1]; p.sz = fo->hi[2] - fo->lo[2];
  } else {          // hi-lo+1 voxels: one more than the volume can supply when hi == dim (src/MarchingCubes.c:1088-1090)
    p.sx = fo->hi[0] - fo->lo[0] + 1; p.sy = fo->hi[1] - fo->lo[1] + 1; p.sz = fo->hi[2] - fo->lo[2] + 1;
  }
  memset(mesh, 0, sizeof(*mesh));
  mesh->classic_soup = p.classic;
  if (p.sx < 2 || p.sy < 2 || p.sz < 2) return B2M_FAIL;
  {  // owned sub-volume planes
    int a = sl.z0 - p.lo2, b = sl.z0 + sl.nzl - p.lo2;
    if (a < 0) a = 0;
    if (sl.z0 + sl.nzl == sl.gnz || b > p.sz) b = p.sz;
    p.zs0 = a < p.sz ? a : p.sz;
    p.zn = b > p.zs0 ? b - p.zs0 : 0;
  }
  p.segs = (p.sx + 31) / 32;
  p.tab = b2m_ptr<signed char>(ctx, BUF_TABLES);
  p.lutinfo = reinterpret_cast<const uint32_t *>(b2m_ptr<char>(ctx, BUF_TABLES) + MCT_INFO_OFF) +
              ((p.classic || p.original_mc) ? 256 : 0);
  p.sc = d_sc;
  // pad value: (min of the composed volume) - iso (src/MarchingCubes.c:1097-1102).  It is only ever
  // read when hi == dim on some axis, and it only matters when a face voxel can be bright, i.e.
  // when edge_max >= iso; otherwise any "outside" value gives the same mesh and we skip the
  // extra reduction pass over the volume.
  p.pad = fo->vmin - fo->iso;
  if (!p.classic) {
    bool touched = fo->hi[0] == g.nx || fo->hi[1] == g.ny || fo->hi[2] == sl.gnz;
    const float de = fo->edge_max - fo->iso;
    if (W > 1 && touched && de > -FLT_EPSILON) {
      // the wrapped extra column / row of the reference would reach two planes ahead
      b2m_set_error("slab mode needs darkened faces below the isolevel (edge_max %g >= iso %g)", fo->edge_max, fo->iso);
      return B2M_EARG;
    }
    p.closed = W > 1;
    if (touched && !(fo->edge_max < fo->iso)) {
      if (W > 1) {
        // the wrapped extra column / row of the reference would reach two planes ahead
        b2m_set_error("slab mode needs darkened faces below the isolevel (edge_max %g >= iso %g)", fo->edge_max, fo->iso);
        return B2M_EARG;
      }
      ctx->h_scalars->cmin_enc = 0xffffffffu;
      CU_TRY(cudaMemcpyAsync(&d_sc->cmin_enc, &ctx->h_scalars->cmin_enc, 4, cudaMemcpyHostToDevice, ctx->stream));
      B2M_TRY(b2m_compose_materialize(ctx, g, fo, nullptr, nullptr, d_sc, 1));
      B2M_TRY(b2m_fetch_scalars(ctx));
      p.pad = f32_dec(ctx->h_scalars->cmin_enc) - fo->iso;
    }
  }
  {
    float pv = p.pad;
    if (fabsf(pv) < FLT_EPSILON) pv = FLT_EPSILON;
    p.pad_inside = p.classic ? 0u : (pv > 0.0f ? 1u : 0u);
  }
  const size_t prow = (size_t)p.sy * p.segs;          // segments per sub-volume plane
  const size_t nseg = prow * p.zn;                    // own segments
  const size_t nseg_all = prow * (p.zn + 1);          // + the next rank's first plane
  const size_t nvox = (size_t)p.sy * p.zn * p.sx;
  B2M_TRY(b2m_reserve(ctx, BUF_SEG, nseg_all * sizeof(uint4)));
  B2M_TRY(b2m_reserve(ctx, BUF_SEG2, (nseg + 1) * 3 * sizeof(uint32_t)));
  p.segbits = b2m_ptr<uint4>(ctx, BUF_SEG);
  p.segt = b2m_ptr<uint32_t>(ctx, BUF_SEG2);
  p.segc = p.segt + nseg;
  p.segcnt = p.segc + nseg;
  size_t cap = nvox / 8 + 65536;
  if (cap > nvox) cap = nvox;
  if (cap < 1) cap = 1;
  unsigned n_active = 0;
  dim3 grid(b2m_cdiv((size_t)p.sy * p.segs, MCB_THREADS), b2m_cdiv(p.zn > 0 ? p.zn : 1, MCB_ZC));
  // the retry decision must be the same on every rank (collectives follow): overflow anywhere -> all redo
  for (int attempt = 0; attempt < 2; attempt++) {
    B2M_TRY(b2m_reserve(ctx, BUF_ACTIVE, cap * sizeof(uint4)));
    p.active = b2m_ptr<uint4>(ctx, BUF_ACTIVE);
    p.active_cap = (unsigned)cap;
    CU_TRY(cudaMemsetAsync(&d_sc->n_active, 0, 4, ctx->stream));
    if (p.zn > 0) KT_LAUNCH(ctx, "mc_classify", k_mc_classify<<<grid, MCB_THREADS, 0, ctx->stream>>>(p));
    CU_TRY(cudaGetLastError());
    B2M_TRY(mc_scan3_totals(ctx, p, nseg, d_sc));
    B2M_TRY(b2m_sync_scalars(ctx, comm));
    n_active = ctx->h_scalars->n_active;
    bool any_over = false;  // every rank takes the same decision: collectives follow
    for (int r = 0; r < W; r++) any_over |= (b2m_sc(ctx, comm, r)->overflow & 1u) != 0;
    if (!any_over) break;
    if (attempt == 1) { b2m_set_error("mc: active list overflow"); return B2M_ECUDA; }
    // rare: more active voxels than the first guess; redo the classification with an exact capacity
    if (n_active > cap) cap = n_active;
    CU_TRY(cudaMemsetAsync(&d_sc->overflow, 0, 4, ctx->stream));
    CU_TRY(cudaMemsetAsync(&d_sc->first_cube, 0xff, 8, ctx->stream));
  }
  unsigned e_off = 0, c_off = 0, t_off = 0;
  unsigned long long NVE = 0, NVC = 0, NT = 0;
  for (int r = 0; r < W; r++) {
    const b2m_scalars *sr = b2m_sc(ctx, comm, r);
    if (r == sl.rank) { e_off = (unsigned)NVE; c_off = (unsigned)NVC; t_off = (unsigned)NT; }
    NVE += sr->tot_v; NVC += sr->tot_c; NT += sr->tot_t;
  }
  const unsigned tot_v = ctx->h_scalars->tot_v, tot_t = ctx->h_scalars->tot_t, tot_c = ctx->h_scalars->tot_c;
  if (NVE + NVC > 0x7fffffffull || NT > 0x7fffffffull || 3ull * NT > 0xffffffffull) {
    b2m_set_error("mesh exceeds the int counts of the meshify() API");
    return B2M_EARG;
  }
  mesh->nv_edge = tot_v; mesh->nv_c = tot_c; mesh->nt = tot_t;
  mesh->NVE = (unsigned)NVE; mesh->NVC = (unsigned)NVC; mesh->NT = (unsigned)NT;
  mesh->e_off = e_off; mesh->c_off = c_off; mesh->t_off = t_off;
  // reference failure rule: < 3 vertices or < 1 triangle (src/MarchingCubes.c:1119, src/oldcubes.c:497)
  if (p.classic ? (3ull * NT < 3) : (NVE + NVC < 3 || NT < 1)) return B2M_FAIL;
  if (ctx->counts_hook) ctx->counts_hook(ctx->hook_user, (size_t)tot_v + tot_c, (size_t)tot_t);
  B2M_TRY(mc_scan3_apply(ctx, p, nseg, e_off, W > 1 && p.zn >= 2 ? prow : ~(size_t)0));
  if (W > 1) {
    // the first own plane of segment records goes down: the rank below needs the vertex numbering of the
    // plane above its last cubes (Lewiner edge codes 4..7, src/MarchingCubes.c:813-825)
    if (p.zn == 0) CU_TRY(cudaMemsetAsync(p.segbits, 0, prow * sizeof(uint4), ctx->stream));
    B2M_TRY(b2m_comm_exchange(ctx, comm, nullptr, 0, nullptr, 0, p.segbits, sl.hl ? prow * sizeof(uint4) : 0,
                              p.segbits + nseg, sl.hh ? prow * sizeof(uint4) : 0));
    // n_first = global vertex base of the second own plane (0xffffffff: the first own plane holds every own
    // vertex): the first-plane vertices are sent down after the emit pass
    if (p.zn >= 2) CU_TRY(cudaMemcpyAsync(&d_sc->n_first, &p.segbits[prow].w, 4, cudaMemcpyDeviceToDevice, ctx->stream));
    else CU_TRY(cudaMemsetAsync(&d_sc->n_first, 0xff, 4, ctx->stream));
  }
  B2M_TRY(b2m_reserve(ctx, BUF_VERTS, ((size_t)tot_v + tot_c) * 24));
  B2M_TRY(b2m_reserve(ctx, BUF_TRIS, (size_t)tot_t * 12));
  const size_t near_words = (size_t)tot_v / 32 + 1;
  B2M_TRY(b2m_reserve(ctx, BUF_REMAP, near_words * 4));
  size_t ccap = ((size_t)tot_v + tot_c) / (p.classic ? 4 : 16) + 4096;
  mc_emit_params e;
  memset(&e, 0, sizeof(e));
  unsigned long long first_cube = ~0ull;
  for (int r = 0; r < W; r++) { const unsigned long long f = b2m_sc(ctx, comm, r)->first_cube; if (f < first_cube) first_cube = f; }
  for (int attempt = 0; attempt < 2; attempt++) {
    B2M_TRY(b2m_reserve(ctx, BUF_CAND, ccap * sizeof(b2m_item)));
    e.verts = b2m_ptr<double>(ctx, BUF_VERTS);
    e.tris = b2m_ptr<int>(ctx, BUF_TRIS);
    e.items = b2m_ptr<b2m_item>(ctx, BUF_CAND);
    e.nearbits = b2m_ptr<uint32_t>(ctx, BUF_REMAP);
    CU_TRY(cudaMemsetAsync(e.nearbits, 0, near_words * 4, ctx->stream));
    e.item_cap = (unsigned)ccap;
    e.n_active = n_active;
    e.e_off = e_off;
    e.nv_edge_l = tot_v;
    e.c_base = (unsigned)NVE + c_off;
    e.t_off = t_off;
    e.first_cube = first_cube;
    CU_TRY(cudaMemsetAsync(&d_sc->n_cand, 0, 4, ctx->stream));
    if (n_active) {
      if (p.classic) KT_LAUNCH(ctx, "mc_emit", k_mc_emit<true><<<b2m_cdiv(n_active, 128), 128, 0, ctx->stream>>>(p, e));
      else KT_LAUNCH(ctx, "mc_emit", k_mc_emit<false><<<b2m_cdiv(n_active, 128), 128, 0, ctx->stream>>>(p, e));
    }
    CU_TRY(cudaGetLastError());
    if (W > 1 && !p.classic && tot_v + tot_c > 0)
      CU_TRY(cudaMemcpyAsync(d_sc->v0, e.verts, 24, cudaMemcpyDeviceToDevice, ctx->stream));
    B2M_TRY(b2m_sync_scalars(ctx, comm));
    bool any_over = false;  // same decision on every rank
    for (int r = 0; r < W; r++) any_over |= (b2m_sc(ctx, comm, r)->overflow & 16u) != 0;
    if (!any_over) break;
    if (attempt == 1) { b2m_set_error("mc: item list overflow"); return B2M_ECUDA; }
    if (ctx->h_scalars->n_cand > ccap) ccap = ctx->h_scalars->n_cand;
    CU_TRY(cudaMemsetAsync(&d_sc->overflow, 0, 4, ctx->stream));
  }
  mesh->nitems = ctx->h_scalars->n_cand;
  mesh->verts = b2m_ptr<double>(ctx, BUF_VERTS);
  mesh->tris = b2m_ptr<int>(ctx, BUF_TRIS);
  mesh->nearbits = b2m_ptr<uint32_t>(ctx, BUF_REMAP);
  // key origin of the weld = the reference's pts[0]
  if (W == 1) {
    mesh->d_p0 = p.classic ? d_sc->pts0 : mesh->verts;
  } else {
    double p0[3] = {0, 0, 0};
    if (p.classic) {
      for (int r = 0; r < W; r++) {
        const b2m_scalars *sr = b2m_sc(ctx, comm, r);
        if (sr->first_cube == first_cube && first_cube != ~0ull) { memcpy(p0, sr->pts0, 24); break; }
      }
    } else {
      for (int r = 0; r < W; r++) {
        const b2m_scalars *sr = b2m_sc(ctx, comm, r);
        if (sr->tot_v + sr->tot_c > 0) { memcpy(p0, sr->v0, 24); break; }
      }
    }
    memcpy(ctx->h_scalars->pts0, p0, 24);
    CU_TRY(cudaMemcpyAsync(d_sc->pts0, ctx->h_scalars->pts0, 24, cudaMemcpyHostToDevice, ctx->stream));
    mesh->d_p0 = d_sc->pts0;
    // the next rank's first-plane vertices: the last own cubes reference them
    const unsigned my_first = ctx->h_scalars->n_first != 0xffffffffu ? ctx->h_scalars->n_first - e_off : tot_v;
    unsigned up_first = 0, up_off = 0;
    if (sl.hh) {
      const b2m_scalars *su = b2m_sc(ctx, comm, sl.rank + 1);
      up_off = e_off + tot_v;  // = the upper rank's first vertex id
      up_first = su->n_first != 0xffffffffu ? su->n_first - up_off : su->tot_v;
    }
    B2M_TRY(b2m_reserve(ctx, BUF_HALO_V, (size_t)up_first * 24 + 24));
    B2M_TRY(b2m_comm_exchange(ctx, comm, nullptr, 0, nullptr, 0, mesh->verts, sl.hl ? (size_t)my_first * 24 : 0,
                              ctx->buf[BUF_HALO_V].p, sl.hh ? (size_t)up_first * 24 : 0));
    mesh->halo_verts = b2m_ptr<double>(ctx, BUF_HALO_V);
    mesh->halo0 = up_off;
    mesh->halo1 = up_off + up_first;
  }
  return B2M_OK;
}
