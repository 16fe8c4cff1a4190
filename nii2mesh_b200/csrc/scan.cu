// scan.cu — device-wide primitives written for this library: exclusive prefix sum (u32) and an
// LSD radix sort of u64 keys (the "GPU radix sort" of the weld; replaces the reference's
// src/radixsort.c:277-348 three-pass 11/11/10-bit CPU sort, here 8-bit digits over a composite
// (f32 key bits << 32 | original index) so the order equals a stable sort by key).
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (unsigned)d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total, uint32_t *sm /* >= 33 */) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  uint32_t inc = warp_incl_scan(v);
  if (lane == 31) sm[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nwarp ? sm[lane] : 0;
    uint32_t winc = warp_incl_scan(w);
    sm[lane] = winc - w;
    if (lane == 31) sm[32] = winc;
  }
  __syncthreads();
  uint32_t res = inc - v + sm[warp];
  *total = sm[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t *__restrict__ in, size_t n,
                                                              uint32_t *__restrict__ partial) {
  __shared__ uint32_t sm[33];
  size_t base = (size_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  uint32_t tot;
  block_excl_scan(s, &tot, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// single block: exclusive scan of `n` partials in place, total -> *total (may be null)
__global__ void __launch_bounds__(1024) k_scan_single(uint32_t *p, size_t n, uint32_t *total) {
  __shared__ uint32_t sm[33];
  uint32_t carry = 0;
  for (size_t base = 0; base < n; base += blockDim.x) {
    size_t idx = base + threadIdx.x;
    uint32_t v = idx < n ? p[idx] : 0;
    uint32_t tot;
    uint32_t ex = block_excl_scan(v, &tot, sm);
    if (idx < n) p[idx] = ex + carry;
    carry += tot;
  }
  if (total && threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t *in, uint32_t *out, size_t n,
                                                             const uint32_t *__restrict__ partial) {
  __shared__ uint32_t sm[33];
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    size_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan(s, &tot, sm) + partial[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    size_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
}

static int scan_rec(b2m_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *d_total, int level) {
  if (n == 0) {
    if (d_total) CU_TRY(cudaMemsetAsync(d_total, 0, 4, ctx->stream));
    return B2M_OK;
  }
  if (n <= 32768 && d_in == d_out) {  // one block, a few trips: cheaper than three launches (radix-sort histograms of small lists)
    KT_LAUNCH(ctx, "scan_single", k_scan_single<<<1, 1024, 0, ctx->stream>>>(d_out, n, d_total));
    CU_TRY(cudaGetLastError());
    return B2M_OK;
  }
  size_t nblk = (n + SCAN_TILE - 1) / SCAN_TILE;
  int which = level == 0 ? BUF_SCAN1 : BUF_SCAN2;
  if (level > 1) {
    b2m_set_error("scan: input too large");
    return B2M_EARG;
  }
  B2M_TRY(b2m_reserve(ctx, which, nblk * 4));
  uint32_t *part = b2m_ptr<uint32_t>(ctx, which);
  KT_LAUNCH(ctx, "scan_reduce", k_scan_reduce<<<(unsigned)nblk, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, part));
  if (nblk <= 65536) {
    KT_LAUNCH(ctx, "scan_single", k_scan_single<<<1, 1024, 0, ctx->stream>>>(part, nblk, d_total));
  } else {
    B2M_TRY(scan_rec(ctx, part, part, nblk, d_total, level + 1));
  }
  KT_LAUNCH(ctx, "scan_apply", k_scan_apply<<<(unsigned)nblk, SCAN_THREADS, 0, ctx->stream>>>(d_in, d_out, n, part));
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

int b2m_exclusive_scan_u32(b2m_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, size_t n, uint32_t *d_total) {
  return scan_rec(ctx, d_in, d_out, n, d_total, 0);
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, u64 keys, ascending.  Stable per pass.
#define RS_THREADS 256
#define RS_ROUNDS 8
#define RS_TILE (RS_THREADS * RS_ROUNDS)

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t *__restrict__ keys, size_t n, int shift,
                                                        uint32_t *__restrict__ hist, unsigned nblk) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ROUNDS; r++) {
    size_t idx = base + (size_t)r * RS_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&h[(unsigned)(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t *__restrict__ keys,
                                                           uint64_t *__restrict__ out, size_t n, int shift,
                                                           const uint32_t *__restrict__ hist, unsigned nblk) {
  __shared__ uint32_t base[256];
  __shared__ uint32_t cnt[RS_THREADS / 32][256];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  base[threadIdx.x] = hist[(size_t)threadIdx.x * nblk + blockIdx.x];
  size_t tile = (size_t)blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ROUNDS; r++) {
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; w++) cnt[w][threadIdx.x] = 0;
    __syncthreads();
    size_t idx = tile + (size_t)r * RS_THREADS + threadIdx.x;
    bool valid = idx < n;
    uint64_t k = valid ? keys[idx] : 0;
    unsigned d = valid ? ((unsigned)(k >> shift) & 255u) : 0xffffffffu;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    unsigned rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) cnt[warp][d] = __popc(peers);
    __syncthreads();
    // digit = threadIdx.x: exclusive prefix over warps, then advance the running base
    {
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < RS_THREADS / 32; w++) {
        uint32_t c = cnt[w][threadIdx.x];
        cnt[w][threadIdx.x] = run;
        run += c;
      }
      __syncthreads();
      if (valid) out[(size_t)base[d] + cnt[warp][d] + rank] = k;
      __syncthreads();
      base[threadIdx.x] += run;
    }
    __syncthreads();
  }
}

int b2m_sort_u64(b2m_ctx *ctx, uint64_t *d_keys, size_t n, int key_bits) {
  if (n < 2) return B2M_OK;
  unsigned nblk = (unsigned)((n + RS_TILE - 1) / RS_TILE);
  B2M_TRY(b2m_reserve(ctx, BUF_SORTB, n * 8));
  B2M_TRY(b2m_reserve(ctx, BUF_SORTH, (size_t)nblk * 256 * 4));
  uint64_t *a = d_keys, *b = b2m_ptr<uint64_t>(ctx, BUF_SORTB);
  uint32_t *hist = b2m_ptr<uint32_t>(ctx, BUF_SORTH);
  int passes = (key_bits + 7) / 8;
  if (passes & 1) passes++;  // even number of passes: the result ends in d_keys
  for (int p = 0; p < passes; p++) {
    int shift = 8 * p;
    KT_LAUNCH(ctx, "rs_hist", k_rs_hist<<<nblk, RS_THREADS, 0, ctx->stream>>>(a, n, shift, hist, nblk));
    B2M_TRY(b2m_exclusive_scan_u32(ctx, hist, hist, (size_t)nblk * 256, nullptr));
    KT_LAUNCH(ctx, "rs_scatter", k_rs_scatter<<<nblk, RS_THREADS, 0, ctx->stream>>>(a, b, n, shift, hist, nblk));
    uint64_t *t = a; a = b; b = t;
  }
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}
