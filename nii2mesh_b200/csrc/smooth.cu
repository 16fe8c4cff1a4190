// smooth.cu — kernel 1: separable 5-tap pre-smooth, fused x->y->z in one pass over the volume,
// plus the intensity-range reduction and the threshold-to-bit-rows kernel.
//
// Reference behaviour reproduced bit-for-bit (/root/reference/src/meshify.c:170-216, quick_smooth):
//   out = (a*k2)+(b*k1)+(c*k0)+(d*k1)+(e*k2), k = 0.45/0.225/0.05 as C doubles, summed left to
//   right in FP64 (no FMA: __dmul_rn/__dadd_rn), ONE rounding to f32 per pass; voxels whose index
//   on the pass axis is < 2 or >= n-2 keep the previous pass's value; nothing happens when a
//   dim < 5 (the reference ignores quick_smooth's EXIT_FAILURE, meshify.c:301).
#include <cuda.h>

#include "common.cuh"

// ---- tile geometry ----
// CTA = 12 warps on a 128 x 24 xy tile (28 staged rows).  Every warp owns TWO adjacent output rows (8 voxels per lane):
// the per-plane fixed work of a warp (addressing, input screen, barrier, loop) is paid once per 8 voxels, and the
// y pass reads 6 staged rows for 2 output rows instead of 5 for 1.  The 28 staged rows are split 2 per warp plus one
// more for warps 0..3, which sit on the four different schedulers: the x-pass work is balanced across them.
#define SX_TX 128                 /* tile width: 32 lanes x 4 voxels */
#ifndef SX_TY
#define SX_TY 24                  /* output rows per tile */
#endif
#define SX_ROWS (SX_TY + 4)       /* rows staged per plane (2 + 2 halo) */
#ifndef SX_STAGE_WARPS
#define SX_STAGE_WARPS 1          /* 1: two more warps stage the last four rows (and own no output rows) instead of a third row in warps 0..3 */
#endif
#if SX_STAGE_WARPS
#define SX_WARPS (SX_ROWS / 2)
#else
#define SX_WARPS (SX_TY / 2)
#endif
#define SX_XTRA (SX_ROWS - 2 * SX_WARPS)  /* staged rows beyond two per warp */
#define SX_THREADS (32 * SX_WARPS)
#define SX_RING 4                 /* planes of x-pass results in shared memory */
#define SX_SMEM (SX_RING * SX_ROWS * SX_TX * 8)
/* 448 threads get 128 registers each: the register file is handed out as if the CTA had 16 warps (__maxnreg__(144)
   compiles without spills and then fails to launch: "too many resources requested") */
#define SX_BOUNDS __launch_bounds__(SX_THREADS, 1)

// mbarrier (shared-memory arrive/wait barrier with phases): split arrive / wait lets warps run one plane apart
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nMB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra MB_DONE;\nbra MB_WAIT;\nMB_DONE:\n}" ::"r"(
          (unsigned)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

#define K0 0.45
#define K1 0.225
#define K2 0.05

__device__ __forceinline__ double fir5(double a, double b, double c, double d, double e) {
  double s = __dmul_rn(a, K2);
  s = __dadd_rn(s, __dmul_rn(b, K1));
  s = __dadd_rn(s, __dmul_rn(c, K0));
  s = __dadd_rn(s, __dmul_rn(d, K1));
  s = __dadd_rn(s, __dmul_rn(e, K2));
  return s;
}

// (double)(float)d without the two 1/16-rate F2F conversions (measured on B200: DADD/DMUL issue at
// 64/clk/SM, F2F.F64<->F32 at 16/clk/SM): adding 1.5*2^(e+29) makes the FP64 adder round d to the
// f32 grid of its binade (round-to-nearest-even, one rounding); subtracting it back is exact.
// FAST is exact when d is +0 or 2^-126 <= |d| < 2^127, which holds for every intermediate of the
// filter when all its inputs are +0 or have 2^-100 <= |v| < 2^100 (sums of products with the
// positive taps cannot fall below 2^-53 of their largest term unless they cancel to +0 exactly).
// Inputs are screened when they are loaded (input_unsafe); a tile that sees anything else
// (denormals, -0, huge values, inf, nan) takes the real conversions for that plane.
template <bool FAST>
__device__ __forceinline__ double round_to_f32(double d) {
  if (FAST) {
    const int eh = __double2hiint(d) & 0x7ff00000;
    const double m = __hiloint2double(eh + 0x01d80000, 0);
    return __dsub_rn(__dadd_rn(d, m), m);
  }
  return (double)(float)d;
}
__device__ __forceinline__ unsigned input_bias(float f) {  // < 0x64000000 iff 2^-100 <= |f| < 2^100
  return (__float_as_uint(f) & 0x7fffffffu) - 0x0d800000u;
}
__device__ __forceinline__ bool input_unsafe(float f) {
  const unsigned u = __float_as_uint(f);
  return u != 0u && ((u & 0x7fffffffu) - 0x0d800000u) >= 0x64000000u;  // not +0 and outside [2^-100, 2^100)
}

template <bool VEC>
__device__ __forceinline__ void load_row4(const float *__restrict__ row, int gx, int nx, float f[4]) {
  if (VEC) {
    if (gx < nx) {
      float4 t = __ldg(reinterpret_cast<const float4 *>(row + gx));
      f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (gx + k < nx) f[k] = __ldg(row + gx + k);
  }
}

// x pass of one staged row: 4 outputs per lane from the lane's float4 and its neighbours' (shuffled as
// floats), rounded to the f32 grid, parked in shared memory as doubles
// XEDGE: the tile touches x < 2 or x >= nx-2 (block-uniform), only then the pass-through selects are needed
template <bool FAST, bool XEDGE>
__device__ __forceinline__ void x_pass_row(const float raw[4], const float hal[2], int lane, int gx, int nx,
                                           double2 *__restrict__ dst) {
  float f[8];
#pragma unroll
  for (int k = 0; k < 4; k++) f[2 + k] = raw[k];
  f[0] = __shfl_up_sync(0xffffffffu, raw[2], 1);
  f[1] = __shfl_up_sync(0xffffffffu, raw[3], 1);
  f[6] = __shfl_down_sync(0xffffffffu, raw[0], 1);
  f[7] = __shfl_down_sync(0xffffffffu, raw[1], 1);
  if (lane == 0) { f[0] = hal[0]; f[1] = hal[1]; }
  if (lane == 31) { f[6] = hal[0]; f[7] = hal[1]; }
  double v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = (double)f[k];
  double o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = gx + k;
    const double r = round_to_f32<FAST>(fir5(v[k], v[k + 1], v[k + 2], v[k + 3], v[k + 4]));
    o[k] = (XEDGE && (x < 2 || x >= nx - 2)) ? v[k + 2] : r;
  }
  dst[lane] = make_double2(o[0], o[1]);
  dst[32 + lane] = make_double2(o[2], o[3]);
}

// y pass (2 adjacent rows x 4 voxels per lane out of 6 staged rows) + z streaming accumulators
// YEDGE: the tile holds rows y < 2 or y >= ny-2 (block-uniform); yb0 / yb1 say which of the two rows pass through
// nobody in the kernel reads what this writes: no "memory" clobber, so it does not pin the loads around it
__device__ __forceinline__ void st_global_if(unsigned *p, unsigned v, bool on) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.global.u32 [%0], %1;\n}" ::"l"(p), "r"(v), "r"((unsigned)on));
}
struct yz_no_hook {
  __device__ __forceinline__ void operator()() const {}
};
// hook(): independent work of the caller placed INSIDE this straight-line block (after the first shared-memory loads
// have been issued) so that the compiler can interleave it with the FP64 chains - see the bit rows of k_smooth3
template <bool YEDGE, class HOOK = yz_no_hook>
__device__ __forceinline__ void yz_pass2(const double2 *__restrict__ buf, int r0, int lane, bool yb0, bool yb1,
                                         double S[2][4][4], float ob[2][4], float oi[2][4], HOOK hook = HOOK()) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    double c[6][2];
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const double2 pj = buf[(r0 + j) * 64 + h * 32 + lane];
      c[j][0] = pj.x; c[j][1] = pj.y;
    }
    if (h == 0) hook();
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
      const int k = 2 * h + kk;
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const double f = fir5(c[r][kk], c[r + 1][kk], c[r + 2][kk], c[r + 3][kk], c[r + 4][kk]);
        const bool yb = YEDGE && (r ? yb1 : yb0);
        // the y pass rounds through the real conversions: the f32 value is needed anyway for border planes, and this
        // keeps the tile free of a CTA-wide "unsafe input" flag (only the x pass, warp-local, uses the adder rounding;
        // measured: the adder rounding here was 4 % slower)
        float yf = (float)f;
        if (yb) yf = (float)c[r + 2][kk];
        const double ys = (double)yf;
        ob[r][k] = yf;
        const double q0 = __dmul_rn(ys, K0), q1 = __dmul_rn(ys, K1), q2 = __dmul_rn(ys, K2);
        const double fin = __dadd_rn(S[r][k][3], q2);
        S[r][k][3] = __dadd_rn(S[r][k][2], q1);
        S[r][k][2] = __dadd_rn(S[r][k][1], q0);
        S[r][k][1] = __dadd_rn(S[r][k][0], q1);
        S[r][k][0] = q2;
        oi[r][k] = (float)fin;
      }
    }
  }
}

// One CTA (12 warps): a 128 x 24 xy tile, marching along z over [z0, z1) (+2 halo planes each side).
//   x pass: warp w filters staged rows 2w, 2w+1 (and 24+w for w < 4; staged row r is gy = y0-2+r) in registers
//           (float4 per lane, neighbours by shuffle), rounds to the f32 grid and parks the rows in shared memory as
//           doubles;
//   y pass: warp w owns output rows y0+2w, y0+2w+1: 2 x 4 voxels per lane from 6 staged rows in shared memory;
//   z pass: streaming accumulators - the reference's left-to-right sum
//           ((((a*k2)+b*k1)+c*k0)+d*k1)+e*k2 is advanced by one term per arriving plane, so a
//           column keeps 4 partial sums instead of a 5-plane ring.
// Every voxel is read once (plus tile halos that hit L2) and written once: 8 B/voxel.
// Slabs: the raw input comes in up to three pieces (halo planes from the rank below, own planes, halo
// planes from the rank above); all z in the kernel are GLOBAL plane numbers.
__device__ __forceinline__ const float *smooth_plane(const smooth_src &s, int zg, size_t nxy) {
  int q = zg - s.rz0;
  if (q < s.n_lo) return s.lo + (size_t)q * nxy;
  q -= s.n_lo;
  if (q < s.n_main) return s.main + (size_t)q * nxy;
  return s.hi + (size_t)(q - s.n_main) * nxy;
}

template <bool FAST, bool XEDGE>
__device__ __forceinline__ void x_pass_rows(const float raw[3][4], const float hal[3][2], bool has3, int warp, int lane,
                                            int gx, int nx, double2 *__restrict__ buf) {
  x_pass_row<FAST, XEDGE>(raw[0], hal[0], lane, gx, nx, buf + (2 * warp) * 64);
  x_pass_row<FAST, XEDGE>(raw[1], hal[1], lane, gx, nx, buf + (2 * warp + 1) * 64);
  if (has3) x_pass_row<FAST, XEDGE>(raw[2], hal[2], lane, gx, nx, buf + (2 * SX_WARPS + warp) * 64);
}

// BITS: the kernel also writes the threshold bit rows of its output planes (what k_threshold would compute from them:
// fg = v >= iso, bg = ~fg, mb = the marching-cubes comparison) for the isolevel the caller ASKED for - the 4 GB re-read
// of the smoothed volume by k_threshold goes away whenever that isolevel survives the range check, which needs the
// min/max this very kernel produces (src/meshify.c:316-319); if it does not survive, k_threshold runs as before.
// Needs nx % 32 == 0 (a bit word is then wholly inside or wholly outside the volume) and the vector path.
template <bool VEC, bool BITS>
__global__ void SX_BOUNDS k_smooth3(const __grid_constant__ smooth_src src, float *__restrict__ out, int nx,
                                                           int ny, int zc, unsigned int *__restrict__ mm_enc,
                                                           const smooth_bits sb) {
  extern __shared__ double2 xs2[];  // [SX_RING][SX_ROWS][2][32]
  __shared__ float red[2][SX_WARPS];
  __shared__ __align__(8) unsigned long long mbar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = BITS ? __shfl_sync(0xffffffffu, tid >> 5, 0) : tid >> 5;  // BITS: known to be warp-uniform
  const int x0 = blockIdx.x * SX_TX, y0 = blockIdx.y * SX_TY;
  const int nz = src.gnz;
  const int z0 = src.oz0 + blockIdx.z * zc, z1 = min(z0 + zc, src.oz0 + src.onz);
  const int zs = max(z0 - 2, 0), ze = min(z1 + 2, nz);
  const int gx = x0 + lane * 4;
  const size_t nxy = (size_t)nx * ny;
  double S[2][4][4];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int q = 0; q < 4; q++) S[r][k][q] = 0.0;
  float vmin = INFINITY, vmax = -INFINITY;

  // loop-invariant addressing.  Staged rows of this warp: slot q < 2 -> staged row 2*warp + q, slot 2 -> staged row
  // 2*SX_WARPS + warp (warps 0..SX_XTRA-1 only); row offsets inside a plane fit 31 bits (dims <= 32767)
  const bool has3 = warp < SX_XTRA;
  bool rowok[3];
  int roff[3];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    const int sy = y0 - 2 + (q < 2 ? 2 * warp + q : 2 * SX_WARPS + warp);
    rowok[q] = sy >= 0 && sy < ny && (q < 2 || has3);
    roff[q] = rowok[q] ? sy * nx : 0;
  }
  const int oy0 = y0 + 2 * warp;  // output rows oy0, oy0 + 1
  const bool yb0 = oy0 < 2 || oy0 >= ny - 2, yb1 = oy0 + 1 < 2 || oy0 + 1 >= ny - 2;
  const bool ok0 = oy0 < ny && gx < nx, ok1 = oy0 + 1 < ny && gx < nx;
  const bool xedge = x0 == 0 || x0 + SX_TX > nx - 2;  // block-uniform
  const bool yedge = y0 == 0 || y0 + SX_TY > ny - 2;  // block-uniform
  float *outp = out + (size_t)oy0 * nx + gx;  // + (z - oz0) * nxy (+ nx for the second row)
  // ---- BITS: threshold bit rows --------------------------------------------------------------------------------
  // Lanes 8j .. 8j+7 hold the 32 voxels of word j of a tile row, two rows per warp.  Each lane turns its 4 + 4 voxels
  // into two nibbles; three exchanges (xor 4 - which also hands row 0 to lanes 8j..8j+3 and row 1 to lanes 8j+4..8j+7 -
  // then xor 1, xor 2) leave word j of row 0 in lane 8j and of row 1 in lane 8j+4, which store them.
  // The exchanges are latency this kernel's 3.5 warps per scheduler cannot hide at the end of a trip and instructions
  // it cannot afford (it is issue-bound), so: (1) the nibbles are PARKED and gathered + stored one trip later, inside
  // the straight-line block of the next y/z pass (flush_bits as the hook of yz_pass2); (2) mb is not computed per
  // voxel: it follows from fg (Lewiner: the same bit; classic: the opposite bit) unless some voxel of the warp's rows
  // is within FLT_EPSILON of the isolevel or a NaN, which one screening compare per voxel and a vote detect - only
  // then are the exact mb bits computed and gathered (park_bits, cold path).  Word offsets are 32-bit (a slab has
  // fewer than 2^28 words).
  const int bw = nx >> 5;                                   // words per row (BITS: nx % 32 == 0)
  const int bpw = ny * bw;                                  // words per plane
  const bool brow1 = (lane & 4) != 0;                       // this lane ends up with row 1
  const int bsh = (lane & 7) * 4;
  const int bidx_lane = (oy0 + (brow1 ? 1 : 0)) * bw + (x0 >> 5) + (lane >> 3);  // + (z - oz0) * bpw
  const bool bst_lane = (lane & 3) == 0 && (brow1 ? ok1 : ok0);
  auto gather2 = [&](unsigned a0, unsigned a1) {  // nibbles of row 0 / row 1 -> the word of this lane's row
    const unsigned mine = (brow1 ? a1 : a0) << bsh, other = (brow1 ? a0 : a1) << bsh;
    unsigned x = mine | __shfl_xor_sync(0xffffffffu, other, 4);
    x |= __shfl_xor_sync(0xffffffffu, x, 1);
    x |= __shfl_xor_sync(0xffffffffu, x, 2);
    return x;
  };
  auto fg_nibble = [&](const float v[4]) {
    return (v[0] >= sb.iso ? 1u : 0u) | (v[1] >= sb.iso ? 2u : 0u) | (v[2] >= sb.iso ? 4u : 0u) | (v[3] >= sb.iso ? 8u : 0u);
  };
  auto mb_nibble = [&](const float v[4]) {
    unsigned m = 0u;
#pragma unroll
    for (int k = 0; k < 4; k++)
      m |= ((sb.classic ? v[k] < sb.iso : __fsub_rn(v[k], sb.iso) > -FLT_EPSILON) ? 1u : 0u) << k;
    return m;
  };
  unsigned pf0 = 0u, pf1 = 0u;     // parked fg nibbles of the two rows
  int pidx = 0;                    // word offset of this lane's parked word
  bool pst = false, pmb = false;   // this lane stores the parked word; its mb word follows from fg
  auto flush_bits = [&]() {  // branch-free: every lane shuffles (garbage when nothing is parked), the stores are predicated
    const unsigned x = gather2(pf0, pf1);
    // predicated stores written out: as C++ ifs they became branches that cut the y/z block in pieces
    st_global_if(sb.fg + pidx, x, pst);
    st_global_if(sb.bg + pidx, ~x, pst && sb.bg != nullptr);
    st_global_if(sb.mb + pidx, sb.classic ? ~x : x, pst && pmb);
  };
  auto park_bits = [&](const float v0[4], const float v1[4], int plane_off) {
    pf0 = fg_nibble(v0); pf1 = fg_nibble(v1);
    bool near = false;  // !(|v - iso| >= eps): inside the band where mb may differ from fg, or a NaN
#pragma unroll
    for (int k = 0; k < 4; k++) {
      near |= !(fabsf(__fsub_rn(v0[k], sb.iso)) >= FLT_EPSILON);
      near |= !(fabsf(__fsub_rn(v1[k], sb.iso)) >= FLT_EPSILON);
    }
    const bool odd = __any_sync(0xffffffffu, near);
    pidx = plane_off + bidx_lane; pst = bst_lane; pmb = !odd;
    if (odd) {  // cold
      const unsigned m = gather2(mb_nibble(v0), mb_nibble(v1));
      if (bst_lane) sb.mb[pidx] = m;
    }
  };

  // raw rows of the plane being staged, prefetched one plane ahead
  // hal: lane 0 = left halo pair (x0-2, x0-1), lane 31 = right halo pair (gx+4, gx+5), fetched by ONE load per register
  // (two predicated loads into the same register serialise on the first one's return: a full memory latency per plane).
  // values that are never loaded (rows / columns outside the volume, halo slots of inner lanes) stay 1.0 so that they
  // pass the input screen; the plane pointer runs along z and is recomputed only where the source piece changes
  float raw[3][4], hal[3][2];
#pragma unroll
  for (int q = 0; q < 3; q++) {
    raw[q][0] = raw[q][1] = raw[q][2] = raw[q][3] = 1.f;
    hal[q][0] = hal[q][1] = 1.f;
  }
  const int haloff = lane == 0 ? x0 - 2 : gx + 4;
  const bool hal0 = (lane == 0 && x0 > 0) || (lane == 31 && gx + 4 < nx);
  const bool hal1 = (lane == 0 && x0 > 0) || (lane == 31 && gx + 5 < nx);
  const int zb1 = src.rz0 + src.n_lo, zb2 = zb1 + src.n_main;
  const float *planep = nullptr;
  auto fetch = [&](int zp) {
    if (zp >= ze) return;  // block-uniform; the registers keep the last plane's (screened) values
    if (zp == zs || zp == zb1 || zp == zb2) planep = smooth_plane(src, zp, nxy);
    else planep += nxy;
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const float *rowp = planep + roff[q];
      if (rowok[q]) {
        load_row4<VEC>(rowp, gx, nx, raw[q]);
        if (VEC) {  // nx % 4 == 0: both halo voxels exist together and the pair is 8-byte aligned
          if (hal0) {
            const float2 h = __ldg(reinterpret_cast<const float2 *>(rowp + haloff));
            hal[q][0] = h.x; hal[q][1] = h.y;
          }
        } else {
          if (hal0) hal[q][0] = __ldg(rowp + haloff);
          if (hal1) hal[q][1] = __ldg(rowp + haloff + 1);
        }
      }
    }
  };
  // x pass of the prefetched plane into its ring slot (inputs are screened only now, not at fetch time: touching the
  // prefetched registers earlier would stall on loads that are meant to fly across the y/z passes).
  // cheap screen first: one unsigned max over the biased magnitudes says "all inside [2^-100, 2^100)";
  // only warps that hold something else (zeros included) run the exact per-value test
  auto stage = [&](int zp) {
    double2 *buf = xs2 + (size_t)((zp - zs) % SX_RING) * (SX_ROWS * 64);
    bool warp_bad = false;
    {
      unsigned m = 0u;
#pragma unroll
      for (int q = 0; q < 3; q++) {
#pragma unroll
        for (int k = 0; k < 4; k++) m = max(m, input_bias(raw[q][k]));
        m = max(m, input_bias(hal[q][0])); m = max(m, input_bias(hal[q][1]));
      }
      if (__any_sync(0xffffffffu, m >= 0x64000000u)) {
        bool raw_bad = false;
#pragma unroll
        for (int q = 0; q < 3; q++) {
#pragma unroll
          for (int k = 0; k < 4; k++) raw_bad |= input_unsafe(raw[q][k]);
          raw_bad |= input_unsafe(hal[q][0]) | input_unsafe(hal[q][1]);
        }
        warp_bad = __any_sync(0xffffffffu, raw_bad);
      }
    }
    if (xedge) {
      if (!warp_bad) x_pass_rows<true, true>(raw, hal, has3, warp, lane, gx, nx, buf);
      else x_pass_rows<false, true>(raw, hal, has3, warp, lane, gx, nx, buf);
    } else {
      if (!warp_bad) x_pass_rows<true, false>(raw, hal, has3, warp, lane, gx, nx, buf);
      else x_pass_rows<false, false>(raw, hal, has3, warp, lane, gx, nx, buf);
    }
  };
  // Plane pipeline without a CTA-wide barrier per plane: a warp stages plane zp+1, ARRIVES on that plane's mbarrier and
  // only then waits for plane zp (staged one trip earlier by everybody) before its y/z passes - warps drift up to one
  // plane apart instead of meeting at every plane.  Two mbarriers alternate (the previous phase of the one a warp arrives
  // on is the one it waited for a trip ago); with four ring slots the slot being written (zp+1) was last read for plane
  // zp-3, which every warp finished before it arrived for plane zp-1 - and that arrival has been waited for.
  if (tid == 0) { mbar_init(&mbar[0], SX_THREADS); mbar_init(&mbar[1], SX_THREADS); }
  fetch(zs);
  __syncthreads();
  stage(zs);
  mbar_arrive(&mbar[0]);
  fetch(zs + 1);
  long long zoff = (long long)(zs - src.oz0) * (long long)nxy;  // may start negative: only used for planes inside [oz0, oz0+onz)
  for (int zp = zs; zp < ze; zp++, zoff += (long long)nxy) {
    const int i = zp - zs;
    if (zp + 1 < ze) {  // block-uniform
      stage(zp + 1);
      mbar_arrive(&mbar[(i + 1) & 1]);
      fetch(zp + 2);  // loads fly while this plane's y/z passes run
    }
    mbar_wait(&mbar[i & 1], (unsigned)(i >> 1) & 1u);
    const double2 *buf = xs2 + (size_t)(i % SX_RING) * (SX_ROWS * 64);
    // ---- y pass + z pass ----
    if (!SX_STAGE_WARPS || warp < SX_TY / 2) {
      float ob[2][4], oi[2][4];
      if (!BITS) {
        if (!yedge) yz_pass2<false>(buf, 2 * warp, lane, false, false, S, ob, oi);
        else yz_pass2<true>(buf, 2 * warp, lane, yb0, yb1, S, ob, oi);
      } else {
        if (!yedge) yz_pass2<false>(buf, 2 * warp, lane, false, false, S, ob, oi, flush_bits);
        else yz_pass2<true>(buf, 2 * warp, lane, yb0, yb1, S, ob, oi, flush_bits);
        pst = false;
      }
      const bool zborder = zp < 2 || zp >= nz - 2;
      const int zo = zp - 2;
      const bool emit_border = zborder && zp >= z0 && zp < z1;
      const bool emit_inner = zo >= z0 && zo < z1 && zo >= 2 && zo < nz - 2;
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (r ? ok1 : ok0) {
          if (emit_border) {
            float *dst = outp + zoff + (r ? nx : 0);
            if (VEC) *reinterpret_cast<float4 *>(dst) = make_float4(ob[r][0], ob[r][1], ob[r][2], ob[r][3]);
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (VEC || gx + k < nx) {
                if (!VEC) dst[k] = ob[r][k];
                vmin = fminf(vmin, ob[r][k]); vmax = fmaxf(vmax, ob[r][k]);
              }
          }
          if (emit_inner) {
            float *dst = outp + (zoff - 2 * (long long)nxy) + (r ? nx : 0);
            if (VEC) *reinterpret_cast<float4 *>(dst) = make_float4(oi[r][0], oi[r][1], oi[r][2], oi[r][3]);
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (VEC || gx + k < nx) {
                if (!VEC) dst[k] = oi[r][k];
                vmin = fminf(vmin, oi[r][k]); vmax = fmaxf(vmax, oi[r][k]);
              }
          }
        }
      }
      if (BITS) {  // emit_border / emit_inner are block-uniform (the shuffles need the whole warp)
        if (emit_inner) park_bits(oi[0], oi[1], (zp - 2 - src.oz0) * bpw);
        if (emit_border) {  // at most four planes of the volume: parked rows out first, then these the same way
          flush_bits();
          park_bits(ob[0], ob[1], (zp - src.oz0) * bpw);
          flush_bits();
          pst = false;
        }
      }
    }
  }
  if (BITS && (!SX_STAGE_WARPS || warp < SX_TY / 2)) flush_bits();  // the last plane's rows
  // block reduction of the range
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
  }
  if (lane == 0) { red[0][warp] = vmin; red[1][warp] = vmax; }
  __syncthreads();
  if (tid < 32) {
    vmin = tid < SX_WARPS ? red[0][tid] : INFINITY;
    vmax = tid < SX_WARPS ? red[1][tid] : -INFINITY;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    }
    if (tid == 0 && vmin <= vmax) {
      atomicMin(&mm_enc[0], f32_enc(vmin));
      atomicMax(&mm_enc[1], f32_enc(vmax));
    }
  }
}


// ================================================================================================================
// TMA variant (sm_100a): the raw planes are brought into shared memory by the tensor-memory accelerator instead of by
// per-thread loads.  One elected thread issues ONE cp.async.bulk.tensor per plane - a 3-D box {SXT_BW x SX_ROWS x 1}
// at (x0-4, y0-2, plane) of the raw volume, whose bytes complete an mbarrier transaction (SASS: UTMALDG) - three planes
// ahead of the x pass.  Out-of-volume parts of the box (tile halos beyond the faces, partial tiles) arrive as zeros, so
// the kernel has no halo loads, no neighbour shuffles, no row / column predicates and no prefetch registers: the x pass
// reads its 8 inputs per lane from shared memory (4 x LDS.64).  Everything after the x pass is the kernel above.
// Requirements of the tensor maps: nx % 4 == 0 and 16-byte aligned pieces (the same condition as the vector path).
#define SXT_BW 136                            /* box width: 128 + 2 x 4 (the left halo padded to a 16-byte boundary) */
#define SXT_RAW 3                             /* raw planes in flight */
#define SXT_RAW_BYTES (SX_ROWS * SXT_BW * 4)  /* 15232 = 119 x 128: the slots stay 128-byte aligned */
#define SXT_SMEM (SX_SMEM + SXT_RAW * SXT_RAW_BYTES + 128)
struct smooth_tmaps {
  CUtensorMap lo, main, hi;
};
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// x pass of one staged row out of shared memory: the lane's 4 outputs x0 + 4 lane + k need the box columns
// 4 lane + 2 .. 4 lane + 9
template <bool FAST, bool XEDGE>
__device__ __forceinline__ void x_pass_row_smem(const float f[8], int lane, int gx, int nx, double2 *__restrict__ dst) {
  double v[8];
#pragma unroll
  for (int k = 0; k < 8; k++) v[k] = (double)f[k];
  double o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = gx + k;
    const double r = round_to_f32<FAST>(fir5(v[k], v[k + 1], v[k + 2], v[k + 3], v[k + 4]));
    o[k] = (XEDGE && (x < 2 || x >= nx - 2)) ? v[k + 2] : r;
  }
  dst[lane] = make_double2(o[0], o[1]);
  dst[32 + lane] = make_double2(o[2], o[3]);
}

__global__ void __launch_bounds__(SX_THREADS, 1) k_smooth3_tma(const __grid_constant__ smooth_tmaps maps, const __grid_constant__ smooth_src src,
                                                               float *__restrict__ out, int nx, int ny, int zc, unsigned int *__restrict__ mm_enc) {
  // [raw ring: SXT_RAW x SX_ROWS x SXT_BW f32][x-pass ring: SX_RING x SX_ROWS x 2 x 32 double2]; the pointers are derived
  // without an integer round trip so that the compiler keeps them in the shared address space (LDS / STS, not generic)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *rawbuf = reinterpret_cast<float *>(smem_raw);
  double2 *xs2 = reinterpret_cast<double2 *>(smem_raw + SXT_RAW * SXT_RAW_BYTES);
  __shared__ float red[2][SX_WARPS];
  __shared__ __align__(8) unsigned long long mbar[2], rbar[SXT_RAW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * SX_TX, y0 = blockIdx.y * SX_TY;
  const int nz = src.gnz;
  const int z0 = src.oz0 + blockIdx.z * zc, z1 = min(z0 + zc, src.oz0 + src.onz);
  const int zs = max(z0 - 2, 0), ze = min(z1 + 2, nz);
  const int gx = x0 + lane * 4;
  const size_t nxy = (size_t)nx * ny;
  double S[2][4][4];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int q = 0; q < 4; q++) S[r][k][q] = 0.0;
  float vmin = INFINITY, vmax = -INFINITY;
  const int oy0 = y0 + 2 * warp;  // output rows oy0, oy0 + 1 (warps < SX_TY / 2)
  const bool yb0 = oy0 < 2 || oy0 >= ny - 2, yb1 = oy0 + 1 < 2 || oy0 + 1 >= ny - 2;
  const bool ok0 = oy0 < ny && gx < nx, ok1 = oy0 + 1 < ny && gx < nx;
  const bool xedge = x0 == 0 || x0 + SX_TX > nx - 2;  // block-uniform
  const bool yedge = y0 == 0 || y0 + SX_TY > ny - 2;  // block-uniform
  float *outp = out + (size_t)oy0 * nx + gx;
  const int zb1 = src.rz0 + src.n_lo, zb2 = zb1 + src.n_main;

  // one thread feeds the raw ring: plane zp -> slot (zp - zs) % SXT_RAW
  auto issue = [&](int zp) {
    if (zp >= ze) return;
    const int slot = (zp - zs) % SXT_RAW;
    const CUtensorMap *m = zp < zb1 ? &maps.lo : (zp < zb2 ? &maps.main : &maps.hi);
    const int q = zp < zb1 ? zp - src.rz0 : (zp < zb2 ? zp - zb1 : zp - zb2);
    mbar_expect_tx(&rbar[slot], SXT_RAW_BYTES);
    tma_load_3d(rawbuf + (size_t)slot * (SXT_RAW_BYTES / 4), m, x0 - 4, y0 - 2, q, &rbar[slot]);
  };
  // x pass of plane zp (two staged rows per warp) out of the raw ring into the x-pass ring
  auto stage = [&](int zp) {
    const int i = zp - zs;
    mbar_wait(&rbar[i % SXT_RAW], (unsigned)(i / SXT_RAW) & 1u);
    const float *rp = rawbuf + (size_t)(i % SXT_RAW) * (SXT_RAW_BYTES / 4) + (2 * warp) * SXT_BW + lane * 4 + 2;
    double2 *buf = xs2 + (size_t)(i % SX_RING) * (SX_ROWS * 64);
    float f[2][8];
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float2 t = *reinterpret_cast<const float2 *>(rp + q * SXT_BW + 2 * k);
        f[q][2 * k] = t.x; f[q][2 * k + 1] = t.y;
      }
    unsigned m = 0u;
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
      for (int k = 0; k < 8; k++) m = max(m, input_bias(f[q][k]));
    bool warp_bad = false;
    if (__any_sync(0xffffffffu, m >= 0x64000000u)) {
      bool bad = false;
#pragma unroll
      for (int q = 0; q < 2; q++)
#pragma unroll
        for (int k = 0; k < 8; k++) bad |= input_unsafe(f[q][k]);
      warp_bad = __any_sync(0xffffffffu, bad);
    }
#pragma unroll
    for (int q = 0; q < 2; q++) {
      double2 *dst = buf + (2 * warp + q) * 64;
      if (xedge) {
        if (!warp_bad) x_pass_row_smem<true, true>(f[q], lane, gx, nx, dst);
        else x_pass_row_smem<false, true>(f[q], lane, gx, nx, dst);
      } else {
        if (!warp_bad) x_pass_row_smem<true, false>(f[q], lane, gx, nx, dst);
        else x_pass_row_smem<false, false>(f[q], lane, gx, nx, dst);
      }
    }
  };
  if (tid == 0) {
    mbar_init(&mbar[0], SX_THREADS); mbar_init(&mbar[1], SX_THREADS);
    for (int k = 0; k < SXT_RAW; k++) mbar_init(&rbar[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) { issue(zs); issue(zs + 1); issue(zs + 2); }
  stage(zs);
  mbar_arrive(&mbar[0]);
  long long zoff = (long long)(zs - src.oz0) * (long long)nxy;
  for (int zp = zs; zp < ze; zp++, zoff += (long long)nxy) {
    const int i = zp - zs;
    if (zp + 1 < ze) {  // block-uniform
      stage(zp + 1);
      mbar_arrive(&mbar[(i + 1) & 1]);
    }
    mbar_wait(&mbar[i & 1], (unsigned)(i >> 1) & 1u);
    // every warp has finished the x pass of plane zp (it arrived for it): that plane's raw slot is free again
    if (tid == 0) issue(zp + SXT_RAW);
    const double2 *buf = xs2 + (size_t)(i % SX_RING) * (SX_ROWS * 64);
    if (warp < SX_TY / 2) {
      float ob[2][4], oi[2][4];
      if (!yedge) yz_pass2<false>(buf, 2 * warp, lane, false, false, S, ob, oi);
      else yz_pass2<true>(buf, 2 * warp, lane, yb0, yb1, S, ob, oi);
      const bool zborder = zp < 2 || zp >= nz - 2;
      const int zo = zp - 2;
      const bool emit_border = zborder && zp >= z0 && zp < z1;
      const bool emit_inner = zo >= z0 && zo < z1 && zo >= 2 && zo < nz - 2;
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (r ? ok1 : ok0) {
          if (emit_border) {
            *reinterpret_cast<float4 *>(outp + zoff + (r ? nx : 0)) = make_float4(ob[r][0], ob[r][1], ob[r][2], ob[r][3]);
#pragma unroll
            for (int k = 0; k < 4; k++) { vmin = fminf(vmin, ob[r][k]); vmax = fmaxf(vmax, ob[r][k]); }
          }
          if (emit_inner) {
            *reinterpret_cast<float4 *>(outp + (zoff - 2 * (long long)nxy) + (r ? nx : 0)) = make_float4(oi[r][0], oi[r][1], oi[r][2], oi[r][3]);
#pragma unroll
            for (int k = 0; k < 4; k++) { vmin = fminf(vmin, oi[r][k]); vmax = fmaxf(vmax, oi[r][k]); }
          }
        }
      }
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
  }
  if (lane == 0) { red[0][warp] = vmin; red[1][warp] = vmax; }
  __syncthreads();
  if (tid < 32) {
    vmin = tid < SX_WARPS ? red[0][tid] : INFINITY;
    vmax = tid < SX_WARPS ? red[1][tid] : -INFINITY;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    }
    if (tid == 0 && vmin <= vmax) {
      atomicMin(&mm_enc[0], f32_enc(vmin));
      atomicMax(&mm_enc[1], f32_enc(vmax));
    }
  }
}


// ================================================================================================================
// Warp-specialised variant (sm_100a; B2M_SMOOTH_WS=1): the two halves of the kernel above run in DIFFERENT warps with
// different register budgets (`setmaxnreg`), so that 20 warps fit where the unified kernel holds 14 at 128 registers:
//   warps 0..7   (two warpgroups, shrunk to 56 registers)   x pass: 3.5 staged rows per warp and plane, straight out of
//                the TMA-fed raw ring; warp 0 also issues the tensor loads, two planes ahead
//   warps 8..19  (three warpgroups, grown to 120 registers) y pass + streaming z pass of two output rows each, exactly
//                as above (yz_pass2)
// Hand-over through shared-memory rings with full / free mbarriers per slot (raw planes: 3 slots, x-pass planes: 4).
#define WS_XW 8
#define WS_CW (SX_TY / 2)
#define WS_THREADS ((WS_XW + WS_CW) * 32)
__device__ __forceinline__ void mbar_wait_k(unsigned long long *bar, int use) { mbar_wait(bar, (unsigned)use & 1u); }

__global__ void __launch_bounds__(WS_THREADS, 1) k_smooth3_ws(const __grid_constant__ smooth_tmaps maps, const __grid_constant__ smooth_src src,
                                                              float *__restrict__ out, int nx, int ny, int zc, unsigned int *__restrict__ mm_enc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *rawbuf = reinterpret_cast<float *>(smem_raw);
  double2 *xs2 = reinterpret_cast<double2 *>(smem_raw + SXT_RAW * SXT_RAW_BYTES);
  __shared__ float red[2][WS_CW];
  __shared__ __align__(8) unsigned long long raw_full[SXT_RAW], raw_free[SXT_RAW], xs_full[SX_RING], xs_free[SX_RING];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * SX_TX, y0 = blockIdx.y * SX_TY;
  const int nz = src.gnz;
  const int z0 = src.oz0 + blockIdx.z * zc, z1 = min(z0 + zc, src.oz0 + src.onz);
  const int zs = max(z0 - 2, 0), ze = min(z1 + 2, nz);
  const int np = ze - zs;  // planes this CTA walks
  const int gx = x0 + lane * 4;
  if (tid == 0) {
    for (int k = 0; k < SXT_RAW; k++) { mbar_init(&raw_full[k], 1); mbar_init(&raw_free[k], WS_XW * 32); }
    for (int k = 0; k < SX_RING; k++) { mbar_init(&xs_full[k], WS_XW * 32); mbar_init(&xs_free[k], WS_CW * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float vmin = INFINITY, vmax = -INFINITY;
  if (warp < WS_XW) {
    // ---------------- x-pass warps ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const bool xedge = x0 == 0 || x0 + SX_TX > nx - 2;  // block-uniform
    const int zb1 = src.rz0 + src.n_lo, zb2 = zb1 + src.n_main;
    auto issue = [&](int i) {  // plane zs + i -> raw slot i % SXT_RAW
      const int zp = zs + i, slot = i % SXT_RAW;
      if (i >= SXT_RAW) mbar_wait_k(&raw_free[slot], i / SXT_RAW - 1);  // every x-pass thread has read the previous tenant
      const CUtensorMap *m = zp < zb1 ? &maps.lo : (zp < zb2 ? &maps.main : &maps.hi);
      const int q = zp < zb1 ? zp - src.rz0 : (zp < zb2 ? zp - zb1 : zp - zb2);
      mbar_expect_tx(&raw_full[slot], SXT_RAW_BYTES);
      tma_load_3d(rawbuf + (size_t)slot * (SXT_RAW_BYTES / 4), m, x0 - 4, y0 - 2, q, &raw_full[slot]);
    };
    if (tid == 0) { issue(0); if (np > 1) issue(1); }
    for (int i = 0; i < np; i++) {
      if (tid == 0 && i + 2 < np) issue(i + 2);
      __syncwarp();
      mbar_wait_k(&raw_full[i % SXT_RAW], i / SXT_RAW);
      if (i >= SX_RING) mbar_wait_k(&xs_free[i % SX_RING], i / SX_RING - 1);  // the y/z warps are done with the slot's previous plane
      const float *rp = rawbuf + (size_t)(i % SXT_RAW) * (SXT_RAW_BYTES / 4) + lane * 4 + 2;
      double2 *buf = xs2 + (size_t)(i % SX_RING) * (SX_ROWS * 64);
      for (int r = warp; r < SX_ROWS; r += WS_XW) {  // staged rows warp, warp + 8, ...
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float2 t = *reinterpret_cast<const float2 *>(rp + r * SXT_BW + 2 * k);
          f[2 * k] = t.x; f[2 * k + 1] = t.y;
        }
        unsigned m = 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) m = max(m, input_bias(f[k]));
        bool warp_bad = false;
        if (__any_sync(0xffffffffu, m >= 0x64000000u)) {
          bool bad = false;
#pragma unroll
          for (int k = 0; k < 8; k++) bad |= input_unsafe(f[k]);
          warp_bad = __any_sync(0xffffffffu, bad);
        }
        double2 *dst = buf + r * 64;
        if (xedge) {
          if (!warp_bad) x_pass_row_smem<true, true>(f, lane, gx, nx, dst);
          else x_pass_row_smem<false, true>(f, lane, gx, nx, dst);
        } else {
          if (!warp_bad) x_pass_row_smem<true, false>(f, lane, gx, nx, dst);
          else x_pass_row_smem<false, false>(f, lane, gx, nx, dst);
        }
      }
      mbar_arrive(&raw_free[i % SXT_RAW]);
      mbar_arrive(&xs_full[i % SX_RING]);
    }
  } else {
    // ---------------- y / z-pass warps ----------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    const int cw = warp - WS_XW;  // 0 .. WS_CW-1: output rows y0 + 2 cw, y0 + 2 cw + 1
    const size_t nxy = (size_t)nx * ny;
    double S[2][4][4];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int q = 0; q < 4; q++) S[r][k][q] = 0.0;
    const int oy0 = y0 + 2 * cw;
    const bool yb0 = oy0 < 2 || oy0 >= ny - 2, yb1 = oy0 + 1 < 2 || oy0 + 1 >= ny - 2;
    const bool ok0 = oy0 < ny && gx < nx, ok1 = oy0 + 1 < ny && gx < nx;
    const bool yedge = y0 == 0 || y0 + SX_TY > ny - 2;  // block-uniform
    float *outp = out + (size_t)oy0 * nx + gx;
    long long zoff = (long long)(zs - src.oz0) * (long long)nxy;
    for (int i = 0; i < np; i++, zoff += (long long)nxy) {
      const int zp = zs + i;
      mbar_wait_k(&xs_full[i % SX_RING], i / SX_RING);
      const double2 *buf = xs2 + (size_t)(i % SX_RING) * (SX_ROWS * 64);
      float ob[2][4], oi[2][4];
      if (!yedge) yz_pass2<false>(buf, 2 * cw, lane, false, false, S, ob, oi);
      else yz_pass2<true>(buf, 2 * cw, lane, yb0, yb1, S, ob, oi);
      mbar_arrive(&xs_free[i % SX_RING]);  // (the staged rows are in registers / consumed: yz_pass2 returned their sums)
      const bool zborder = zp < 2 || zp >= nz - 2;
      const int zo = zp - 2;
      const bool emit_border = zborder && zp >= z0 && zp < z1;
      const bool emit_inner = zo >= z0 && zo < z1 && zo >= 2 && zo < nz - 2;
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (r ? ok1 : ok0) {
          if (emit_border) {
            *reinterpret_cast<float4 *>(outp + zoff + (r ? nx : 0)) = make_float4(ob[r][0], ob[r][1], ob[r][2], ob[r][3]);
#pragma unroll
            for (int k = 0; k < 4; k++) { vmin = fminf(vmin, ob[r][k]); vmax = fmaxf(vmax, ob[r][k]); }
          }
          if (emit_inner) {
            *reinterpret_cast<float4 *>(outp + (zoff - 2 * (long long)nxy) + (r ? nx : 0)) = make_float4(oi[r][0], oi[r][1], oi[r][2], oi[r][3]);
#pragma unroll
            for (int k = 0; k < 4; k++) { vmin = fminf(vmin, oi[r][k]); vmax = fmaxf(vmax, oi[r][k]); }
          }
        }
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    }
    if (lane == 0) { red[0][cw] = vmin; red[1][cw] = vmax; }
  }
  __syncthreads();
  if (tid < 32) {
    vmin = tid < WS_CW ? red[0][tid] : INFINITY;
    vmax = tid < WS_CW ? red[1][tid] : -INFINITY;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    }
    if (tid == 0 && vmin <= vmax) {
      atomicMin(&mm_enc[0], f32_enc(vmin));
      atomicMax(&mm_enc[1], f32_enc(vmax));
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn tma_encoder(void) {
  static encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (encode_tiled_fn)p;
    else cudaGetLastError();
    tried = true;
  }
  return fn;
}
// raw planes [nplanes][ny][nx] f32 as a 3-D tensor, box SXT_BW x SX_ROWS x 1, zero fill outside
static bool smooth_make_tmap(CUtensorMap *m, const float *p, int nx, int ny, int nplanes) {
  encode_tiled_fn enc = tma_encoder();
  if (!enc || nplanes < 1) return false;
  const cuuint64_t dim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nplanes};
  const cuuint64_t stride[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
  const cuuint32_t box[3] = {SXT_BW, SX_ROWS, 1}, estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(p), dim, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// intensity range of a volume (reference: src/meshify.c:306-311) when no smoothing precedes it
__global__ void __launch_bounds__(256) k_minmax(const float *__restrict__ in, size_t n, unsigned int *__restrict__ mm_enc) {
  __shared__ float red[2][8];
  float vmin = INFINITY, vmax = -INFINITY;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  // 16-byte vector loads on the aligned middle; the (at most 3) voxels before the first aligned address and the tail go
  // one by one: callers hand in any 4-byte aligned pointer (a slab that starts at plane z0 of a volume whose plane
  // size is not a multiple of 4, a cropped atlas box)
  const size_t head = min(n, (size_t)((16u - (unsigned)(reinterpret_cast<uintptr_t>(in) & 15u)) & 15u) / 4);
  const size_t n4 = (n - head) / 4;
  const float4 *in4 = reinterpret_cast<const float4 *>(in + head);
  for (size_t k = i; k < n4; k += stride) {
    float4 v = __ldg(in4 + k);
    vmin = fminf(fminf(vmin, v.x), fminf(v.y, fminf(v.z, v.w)));
    vmax = fmaxf(fmaxf(vmax, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  const size_t tail0 = head + n4 * 4;  // head and tail hold at most 3 voxels each: threads 0..2 of the grid take them
  if (i < head) { const float v = in[i]; vmin = fminf(vmin, v); vmax = fmaxf(vmax, v); }
  if (tail0 + i < n) { const float v = in[tail0 + i]; vmin = fminf(vmin, v); vmax = fmaxf(vmax, v); }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = vmin; red[1][threadIdx.x >> 5] = vmax; }
  __syncthreads();
  if (threadIdx.x < 32) {
    vmin = threadIdx.x < 8 ? red[0][threadIdx.x] : INFINITY;
    vmax = threadIdx.x < 8 ? red[1][threadIdx.x] : -INFINITY;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
      vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    }
    if (threadIdx.x == 0 && vmin <= vmax) {
      atomicMin(&mm_enc[0], f32_enc(vmin));
      atomicMax(&mm_enc[1], f32_enc(vmax));
    }
  }
}

// mask = img >= iso (src/meshify.c:325-330) as bit rows.  fg word bit b <=> voxel x = 32*xw + b is
// foreground; bg = complement restricted to x < nx.  A warp turns THR_WPW consecutive bit words
// (32 voxels each) per trip: THR_WPW independent 128-byte loads in flight, one ballot each, then
// lanes 0..THR_WPW-1 store the words (the one-word-per-warp first version ran at 1.2 TB/s: too few
// bytes in flight).
#define THR_WPW 8
/* measured on G1024: 0.770 ms at 6 CTAs per SM (40 registers), 0.810 at 5, 0.883 unconstrained (58 registers), 1.02 at 1;
   8 spills the eight values a warp trip holds */
#ifndef THR_MINB
#define THR_MINB 6
#endif
__global__ void __launch_bounds__(256, THR_MINB) k_threshold(const float *__restrict__ in, int nx, int w, long long nwords,
                                                   float iso, uint32_t *__restrict__ fg, uint32_t *__restrict__ bg,
                                                   uint32_t *__restrict__ mb, int classic) {
  const unsigned lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long w0 = warp * THR_WPW; w0 < nwords; w0 += nwarps * THR_WPW) {
    if (nx == w * 32 && w0 + THR_WPW <= nwords) {
      // fast path (whole-word rows, 8 full words): no validity predicates, constant load offsets; per word one
      // compare + ballot for fg and one subtract + compare + ballot for mb
      const float *p = in + w0 * 32 + lane;
      float v[THR_WPW];
#pragma unroll
      for (int j = 0; j < THR_WPW; j++) v[j] = __ldg(p + j * 32);
      uint32_t mf[THR_WPW], mm[THR_WPW];
#pragma unroll
      for (int j = 0; j < THR_WPW; j++) {
        mf[j] = __ballot_sync(0xffffffffu, v[j] >= iso);
        mm[j] = 0u;
        if (mb) mm[j] = classic ? __ballot_sync(0xffffffffu, v[j] < iso) : __ballot_sync(0xffffffffu, __fsub_rn(v[j], iso) > -FLT_EPSILON);
      }
      if (lane == 0) {
        uint4 *f4 = reinterpret_cast<uint4 *>(fg + w0);
        f4[0] = make_uint4(mf[0], mf[1], mf[2], mf[3]); f4[1] = make_uint4(mf[4], mf[5], mf[6], mf[7]);
        if (bg) {
          uint4 *b4 = reinterpret_cast<uint4 *>(bg + w0);
          b4[0] = make_uint4(~mf[0], ~mf[1], ~mf[2], ~mf[3]); b4[1] = make_uint4(~mf[4], ~mf[5], ~mf[6], ~mf[7]);
        }
        if (mb) {
          uint4 *m4 = reinterpret_cast<uint4 *>(mb + w0);
          m4[0] = make_uint4(mm[0], mm[1], mm[2], mm[3]); m4[1] = make_uint4(mm[4], mm[5], mm[6], mm[7]);
        }
      }
      continue;
    }
    float v[THR_WPW];
    bool ok[THR_WPW];
#pragma unroll
    for (int j = 0; j < THR_WPW; j++) {
      const long long word = w0 + j;
      if (nx == w * 32) {  // rows are whole words: voxel index = word * 32 + lane, no division
        ok[j] = word < nwords;
        v[j] = ok[j] ? __ldg(in + word * 32 + lane) : 0.f;
      } else {
        const long long row = word / w;
        const int x = (int)(word - row * w) * 32 + (int)lane;
        ok[j] = word < nwords && x < nx;
        v[j] = ok[j] ? __ldg(in + row * nx + x) : 0.f;
      }
    }
    // every lane gets every ballot, so lane 0 stores the 8 words of each array as two 16-byte vectors (the kernel is
    // ALU-bound - 93 % ALU pipe in ncu - and the per-word "lane == j" selects were 40 % of its ALU work)
    const bool whole = nx == w * 32;
    uint32_t mf[THR_WPW], mv[THR_WPW], mm[THR_WPW];
#pragma unroll
    for (int j = 0; j < THR_WPW; j++) {
      mf[j] = __ballot_sync(0xffffffffu, ok[j] && v[j] >= iso);
      mv[j] = whole ? (ok[j] ? 0xffffffffu : 0u) : __ballot_sync(0xffffffffu, ok[j]);
      mm[j] = 0u;
      if (mb) mm[j] = classic ? __ballot_sync(0xffffffffu, ok[j] && v[j] < iso)
                              : __ballot_sync(0xffffffffu, ok[j] && __fsub_rn(v[j], iso) > -FLT_EPSILON);
    }
    if (lane == 0) {
      if (w0 + THR_WPW <= nwords) {
        uint4 *f4 = reinterpret_cast<uint4 *>(fg + w0);
        f4[0] = make_uint4(mf[0], mf[1], mf[2], mf[3]); f4[1] = make_uint4(mf[4], mf[5], mf[6], mf[7]);
        if (bg) {
          uint4 *b4 = reinterpret_cast<uint4 *>(bg + w0);
          b4[0] = make_uint4(~mf[0] & mv[0], ~mf[1] & mv[1], ~mf[2] & mv[2], ~mf[3] & mv[3]);
          b4[1] = make_uint4(~mf[4] & mv[4], ~mf[5] & mv[5], ~mf[6] & mv[6], ~mf[7] & mv[7]);
        }
        if (mb) {
          uint4 *m4 = reinterpret_cast<uint4 *>(mb + w0);
          m4[0] = make_uint4(mm[0], mm[1], mm[2], mm[3]); m4[1] = make_uint4(mm[4], mm[5], mm[6], mm[7]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < THR_WPW; j++)
          if (w0 + j < nwords) {
            fg[w0 + j] = mf[j];
            if (bg) bg[w0 + j] = ~mf[j] & mv[j];
            if (mb) mb[w0 + j] = mm[j];
          }
      }
    }
  }
}

int b2m_smooth_run(b2m_ctx *ctx, const smooth_src &src, float *d_out, const b2m_geom &g, b2m_scalars *d_sc,
                   const smooth_bits *bits, int *bits_done) {
  if (bits_done) *bits_done = 0;
  const unsigned tx = b2m_cdiv(g.nx, SX_TX), ty = b2m_cdiv(g.ny, SX_TY);
  const int onz = src.onz;
  // z chunks: one CTA per SM at a time, so the step takes ceil(CTAs / SMs) rounds of (zc + 4) planes each; take the
  // split with the cheapest estimate (chunks >= 16 planes keep the 4 halo planes a small overhead)
  int zc = onz;
  {
    const size_t tiles = (size_t)tx * ty;
    size_t best = 0;
    for (int k = 1; k <= 64; k++) {
      const int c = (onz + k - 1) / k;
      if (c < 16 && k > 1) break;
      const size_t ctas = tiles * (size_t)((onz + c - 1) / c);
      const size_t cost = ((ctas + ctx->sm_count - 1) / ctx->sm_count) * (size_t)(c + 4);
      if (!best || cost < best) { best = cost; zc = c; }
    }
  }
  dim3 grid(tx, ty, b2m_cdiv(onz, zc));
  const bool vec = (g.nx % 4 == 0) && (((uintptr_t)src.main | (uintptr_t)src.lo | (uintptr_t)src.hi | (uintptr_t)d_out) % 16 == 0);
  // more than 48 KB of dynamic shared memory: opt in once PER DEVICE (function attributes belong to the device's
  // context; one process may drive several GPUs - local slab groups, atlas threads)
  if (!ctx->smooth_attr_done) {  // per ctx (= per host thread and device): no shared flag to race on
    CU_TRY(cudaFuncSetAttribute(k_smooth3<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SX_SMEM));
    CU_TRY(cudaFuncSetAttribute(k_smooth3<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SX_SMEM));
    CU_TRY(cudaFuncSetAttribute(k_smooth3<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SX_SMEM));
    ctx->smooth_attr_done = 1;
  }
  // TMA path (B2M_SMOOTH_TMA=1; measured on B200 at G1024: 4.16 ms against 3.65 ms for the per-thread loads, so it is
  // not the default): needs what the vector path needs, a stride of the rows that is a multiple of 16 bytes, and a
  // driver that can encode tensor maps
  static const bool want_ws = getenv("B2M_SMOOTH_WS") && atoi(getenv("B2M_SMOOTH_WS")) > 0;
  static const bool want_tma = want_ws || (getenv("B2M_SMOOTH_TMA") && atoi(getenv("B2M_SMOOTH_TMA")) > 0);
  if (vec && want_tma && SX_STAGE_WARPS) {
    smooth_tmaps maps;
    memset(&maps, 0, sizeof(maps));
    bool ok = smooth_make_tmap(&maps.main, src.main, g.nx, g.ny, src.n_main);
    if (ok && src.n_lo) ok = smooth_make_tmap(&maps.lo, src.lo, g.nx, g.ny, src.n_lo);
    if (ok && src.n_hi) ok = smooth_make_tmap(&maps.hi, src.hi, g.nx, g.ny, src.n_hi);
    if (ok) {
      if (!ctx->smooth_tma_attr_done) {
        CU_TRY(cudaFuncSetAttribute(k_smooth3_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, SXT_SMEM));
        CU_TRY(cudaFuncSetAttribute(k_smooth3_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, SXT_SMEM));
        ctx->smooth_tma_attr_done = 1;
      }
      if (want_ws)
        KT_LAUNCH(ctx, "smooth3", k_smooth3_ws<<<grid, WS_THREADS, SXT_SMEM, ctx->stream>>>(maps, src, d_out, g.nx, g.ny, zc, &d_sc->vmin_enc));
      else
        KT_LAUNCH(ctx, "smooth3", k_smooth3_tma<<<grid, SX_THREADS, SXT_SMEM, ctx->stream>>>(maps, src, d_out, g.nx, g.ny, zc, &d_sc->vmin_enc));
      CU_TRY(cudaGetLastError());
      return B2M_OK;
    }
  }
  smooth_bits nob;
  memset(&nob, 0, sizeof(nob));
  if (vec && bits && g.nx % 32 == 0) {
    KT_LAUNCH(ctx, "smooth3", k_smooth3<true, true><<<grid, SX_THREADS, SX_SMEM, ctx->stream>>>(src, d_out, g.nx, g.ny, zc, &d_sc->vmin_enc, *bits));
    if (bits_done) *bits_done = 1;
  } else if (vec)
    KT_LAUNCH(ctx, "smooth3", k_smooth3<true, false><<<grid, SX_THREADS, SX_SMEM, ctx->stream>>>(src, d_out, g.nx, g.ny, zc, &d_sc->vmin_enc, nob));
  else
    KT_LAUNCH(ctx, "smooth3", k_smooth3<false, false><<<grid, SX_THREADS, SX_SMEM, ctx->stream>>>(src, d_out, g.nx, g.ny, zc, &d_sc->vmin_enc, nob));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

int b2m_minmax_run(b2m_ctx *ctx, const float *d_in, size_t n, b2m_scalars *d_sc) {
  unsigned blocks = (unsigned)min((long long)ctx->sm_count * 16, (long long)(n / 4 + 255) / 256 + 1);
  KT_LAUNCH(ctx, "minmax", k_minmax<<<blocks, 256, 0, ctx->stream>>>(d_in, n, &d_sc->vmin_enc));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

int b2m_threshold_run(b2m_ctx *ctx, const float *d_in, const b2m_geom &g, float iso, uint32_t *d_fg, uint32_t *d_bg,
                      uint32_t *d_mb, int classic) {
  long long warps = (g.nwords + THR_WPW - 1) / THR_WPW;
  long long blocks = (warps + 7) / 8;
  const long long cap = (long long)ctx->sm_count * 64;  // grid-stride beyond a few waves
  if (blocks > cap) blocks = cap;
  KT_LAUNCH(ctx, "threshold", k_threshold<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_in, g.nx, g.w, g.nwords, iso, d_fg, d_bg, d_mb, classic));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}
