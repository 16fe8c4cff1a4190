// smooth.cu — kernel 1: separable 5-tap pre-smooth, fused x->y->z in one pass over the volume,
// plus the intensity-range reduction and the threshold-to-bit-rows kernel.
//
// Reference behaviour reproduced bit-for-bit (/root/reference/src/meshify.c:170-216, quick_smooth):
//   out = (a*k2)+(b*k1)+(c*k0)+(d*k1)+(e*k2), k = 0.45/0.225/0.05 as C doubles, summed left to
//   right in FP64 (no FMA: __dmul_rn/__dadd_rn), ONE rounding to f32 per pass; voxels whose index
//   on the pass axis is < 2 or >= n-2 keep the previous pass's value; nothing happens when a
//   dim < 5 (the reference ignores quick_smooth's EXIT_FAILURE, meshify.c:301).
//
// Layout: a CTA owns a 64x16 xy tile and marches along z.  Per plane it stages the raw tile
// (+2 halo) in shared memory as doubles, runs the x pass and the y pass out of shared memory and
// keeps a 5-deep ring of y-smoothed planes per thread in registers for the z pass, so every voxel
// is read from HBM once (halo re-reads hit L2) and written once: 8 B/voxel algorithmic.
#include "common.cuh"

#define SM_TX 64
#define SM_TY 16
#define SM_THREADS 256
#define SM_ZC 64
#define SM_COLS (SM_TX * SM_TY / SM_THREADS) /* 4 outputs per thread per plane */

__device__ __forceinline__ double fir5(double a, double b, double c, double d, double e) {
  const double k0 = 0.45, k1 = 0.225, k2 = 0.05;
  double s = __dmul_rn(a, k2);
  s = __dadd_rn(s, __dmul_rn(b, k1));
  s = __dadd_rn(s, __dmul_rn(c, k0));
  s = __dadd_rn(s, __dmul_rn(d, k1));
  s = __dadd_rn(s, __dmul_rn(e, k2));
  return s;
}

__global__ void __launch_bounds__(SM_THREADS) k_smooth3(const float *__restrict__ in, float *__restrict__ out, int nx,
                                                        int ny, int nz, unsigned int *__restrict__ mm_enc) {
  __shared__ double raw[SM_TY + 4][SM_TX + 4];
  __shared__ double xs[SM_TY + 4][SM_TX];
  __shared__ float red[2][SM_THREADS / 32];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * SM_TX, y0 = blockIdx.y * SM_TY;
  const int z0 = blockIdx.z * SM_ZC;
  const int z1 = min(z0 + SM_ZC, nz);
  const int zs = max(z0 - 2, 0);
  const size_t nxy = (size_t)nx * ny;
  const int lx = tid % SM_TX, lyb = tid / SM_TX;  // lyb in 0..3
  double ring[SM_COLS][5];
#pragma unroll
  for (int c = 0; c < SM_COLS; c++)
#pragma unroll
    for (int k = 0; k < 5; k++) ring[c][k] = 0.0;
  float vmin = INFINITY, vmax = -INFINITY;

  for (int zp = zs; zp < z1 + 2; zp++) {
    if (zp < nz) {
      const float *plane = in + (size_t)zp * nxy;
      for (int i = tid; i < (SM_TY + 4) * (SM_TX + 4); i += SM_THREADS) {
        int ly = i / (SM_TX + 4), lxx = i % (SM_TX + 4);
        int gx = x0 - 2 + lxx, gy = y0 - 2 + ly;
        double v = 0.0;
        if (gx >= 0 && gx < nx && gy >= 0 && gy < ny) v = (double)__ldg(plane + (size_t)gy * nx + gx);
        raw[ly][lxx] = v;
      }
      __syncthreads();
      for (int i = tid; i < (SM_TY + 4) * SM_TX; i += SM_THREADS) {
        int ly = i / SM_TX, lxx = i % SM_TX;
        int gx = x0 + lxx;
        double v;
        if (gx < 2 || gx >= nx - 2) v = raw[ly][lxx + 2];
        else v = (double)(float)fir5(raw[ly][lxx], raw[ly][lxx + 1], raw[ly][lxx + 2], raw[ly][lxx + 3], raw[ly][lxx + 4]);
        xs[ly][lxx] = v;
      }
      __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < SM_COLS; c++) {
      int ly = lyb + (SM_THREADS / SM_TX) * c;
      int gy = y0 + ly;
      double v = 0.0;
      if (zp < nz) {
        if (gy < 2 || gy >= ny - 2) v = xs[ly + 2][lx];
        else v = (double)(float)fir5(xs[ly][lx], xs[ly + 1][lx], xs[ly + 2][lx], xs[ly + 3][lx], xs[ly + 4][lx]);
      }
      ring[c][0] = ring[c][1]; ring[c][1] = ring[c][2]; ring[c][2] = ring[c][3]; ring[c][3] = ring[c][4];
      ring[c][4] = v;
      int zo = zp - 2;
      int gx = x0 + lx;
      if (zo >= z0 && zo < z1 && gx < nx && gy < ny) {
        float o;
        if (zo < 2 || zo >= nz - 2) o = (float)ring[c][2];
        else o = (float)fir5(ring[c][0], ring[c][1], ring[c][2], ring[c][3], ring[c][4]);
        out[(size_t)zo * nxy + (size_t)gy * nx + gx] = o;
        vmin = fminf(vmin, o);
        vmax = fmaxf(vmax, o);
      }
    }
    // raw/xs are rewritten next iteration only after the two barriers above
  }
  // block reduction of the range
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, d));
    vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = vmin; red[1][tid >> 5] = vmax; }
  __syncthreads();
  if (tid < 32) {
    vmin = tid < SM_THREADS / 32 ? red[0][tid] : INFINITY;
    vmax = tid < SM_THREADS / 32 ? red[1][tid] : -INFINITY;
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

// intensity range of a volume (reference: src/meshify.c:306-311) when no smoothing precedes it
__global__ void __launch_bounds__(256) k_minmax(const float *__restrict__ in, size_t n, unsigned int *__restrict__ mm_enc) {
  __shared__ float red[2][8];
  float vmin = INFINITY, vmax = -INFINITY;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t n4 = n / 4;
  const float4 *in4 = reinterpret_cast<const float4 *>(in);
  for (size_t k = i; k < n4; k += stride) {
    float4 v = __ldg(in4 + k);
    vmin = fminf(fminf(vmin, v.x), fminf(v.y, fminf(v.z, v.w)));
    vmax = fmaxf(fmaxf(vmax, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) {
    float v = in[k];
    vmin = fminf(vmin, v);
    vmax = fmaxf(vmax, v);
  }
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

// mask = img >= iso (src/meshify.c:325-330) as bit rows: one warp per 32-voxel word, ballot.
// fg word bit b <=> voxel x = 32*xw + b is foreground; bg = complement restricted to x < nx.
__global__ void __launch_bounds__(256) k_threshold(const float *__restrict__ in, int nx, int w, long long nwords,
                                                   float iso, uint32_t *__restrict__ fg, uint32_t *__restrict__ bg) {
  long long word = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31;
  if (word >= nwords) return;
  long long row = word / w;
  int xw = (int)(word - row * w);
  int x = xw * 32 + (int)lane;
  bool valid = x < nx;
  bool b = valid && (__ldg(in + row * nx + x) >= iso);
  unsigned m = __ballot_sync(0xffffffffu, b);
  unsigned vm = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) {
    fg[word] = m;
    if (bg) bg[word] = ~m & vm;
  }
}

int b2m_smooth_run(b2m_ctx *ctx, const float *d_in, float *d_out, const b2m_geom &g, b2m_scalars *d_sc) {
  dim3 grid(b2m_cdiv(g.nx, SM_TX), b2m_cdiv(g.ny, SM_TY), b2m_cdiv(g.nz, SM_ZC));
  KT_LAUNCH(ctx, "smooth3", k_smooth3<<<grid, SM_THREADS, 0, ctx->stream>>>(d_in, d_out, g.nx, g.ny, g.nz, &d_sc->vmin_enc));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

int b2m_minmax_run(b2m_ctx *ctx, const float *d_in, const b2m_geom &g, b2m_scalars *d_sc) {
  unsigned blocks = (unsigned)min((long long)ctx->sm_count * 16, (g.n / 4 + 255) / 256 + 1);
  KT_LAUNCH(ctx, "minmax", k_minmax<<<blocks, 256, 0, ctx->stream>>>(d_in, (size_t)g.n, &d_sc->vmin_enc));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

int b2m_threshold_run(b2m_ctx *ctx, const float *d_in, const b2m_geom &g, float iso, uint32_t *d_fg, uint32_t *d_bg) {
  long long threads = g.nwords * 32;
  KT_LAUNCH(ctx, "threshold", k_threshold<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_in, g.nx, g.w, g.nwords, iso, d_fg, d_bg));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}
