// isolevel.cu — automatic isolevel selection (-i d / m / b) on a device-resident volume.
//
// Reference (/root/reference/src/isolevel.c): setThreshold() (:245-277) = robust range (2nd..98th percentile of a
// 1001-bin histogram, nifti_robust_range :35-139), a 256-bin histogram over that range (:251-262) and Otsu's method
// on it (nii_otsu :141-243; mode 5 = three thresholds for dark / bright, mode 3 = one for medium).  The reference
// makes three passes over the volume on one core; here the passes are GPU reductions (range + NaN count, two
// histograms with shared-memory privatisation) and only the 256-bin Otsu search runs on the host.
// Bit-exactness: bin index = (int)round((v - mn) * scl) with the subtraction and product in f32 exactly as the C
// expression evaluates them; the Otsu scores are the reference's doubles computed in the reference's order, so ties
// and near-ties pick the same thresholds.
#include <math.h>

#include "common.cuh"

struct iso_range {
  unsigned int mn_enc, mx_enc;
  unsigned long long n_nan, n_zero;
};

__global__ void __launch_bounds__(256) k_iso_range(const float *__restrict__ img, size_t n, iso_range *__restrict__ out) {
  float mn = INFINITY, mx = -INFINITY;
  unsigned long long nn = 0, nz = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = __ldg(img + i);
    if (isnan(v)) { nn++; continue; }
    if (v == 0.0f) nz++;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    nn += __shfl_xor_sync(0xffffffffu, nn, d);
    nz += __shfl_xor_sync(0xffffffffu, nz, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (mn <= mx) { atomicMin(&out->mn_enc, f32_enc(mn)); atomicMax(&out->mx_enc, f32_enc(mx)); }
    if (nn) atomicAdd(&out->n_nan, nn);
    if (nz) atomicAdd(&out->n_zero, nz);
  }
}

// hist[(int)round((v - mn) * scl)]++ for every non-NaN voxel; clamp = the 256-bin variant (isolevel.c:258-260)
#define ISO_MAXBINS 1001
__global__ void __launch_bounds__(256) k_iso_hist(const float *__restrict__ img, size_t n, float mn, float scl, int nbins, int clamp,
                                                  unsigned int *__restrict__ hist) {
  __shared__ unsigned int sh[ISO_MAXBINS];
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = __ldg(img + i);
    if (isnan(v)) continue;
    int idx = (int)round((double)__fmul_rn(__fsub_rn(v, mn), scl));
    if (clamp) { idx = idx < nbins - 1 ? idx : nbins - 1; idx = idx > 0 ? idx : 0; }
    if ((unsigned)idx < (unsigned)nbins) atomicAdd(&sh[idx], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

static int iso_hist(b2m_ctx *ctx, const float *d_img, size_t n, float mn, float scl, int nbins, int clamp, unsigned int *h_hist) {
  B2M_TRY(b2m_reserve(ctx, BUF_ATLAS, ISO_MAXBINS * 4 + 256));
  unsigned int *d_hist = b2m_ptr<unsigned int>(ctx, BUF_ATLAS);
  CU_TRY(cudaMemsetAsync(d_hist, 0, (size_t)nbins * 4, ctx->stream));
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
  KT_LAUNCH(ctx, "iso_hist", k_iso_hist<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_img, n, mn, scl, nbins, clamp, d_hist));
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(h_hist, d_hist, (size_t)nbins * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B2M_OK;
}

// nifti_robust_range(img, nvox, &pct2, &pct98, ignoreZeroVoxels = 0), isolevel.c:35-139
static int robust_range(b2m_ctx *ctx, const float *d_img, size_t nvox, float *pct2, float *pct98) {
  *pct2 = 0.0f;
  *pct98 = 1.0f;
  if (nvox < 1) return 1;
  B2M_TRY(b2m_reserve(ctx, BUF_ATLAS, ISO_MAXBINS * 4 + 256));
  iso_range *d_r = b2m_ptr<iso_range>(ctx, BUF_ATLAS);
  iso_range h;
  h.mn_enc = 0xffffffffu; h.mx_enc = 0u; h.n_nan = 0; h.n_zero = 0;
  CU_TRY(cudaMemcpyAsync(d_r, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  size_t blocks = (nvox + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
  KT_LAUNCH(ctx, "iso_range", k_iso_range<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_img, nvox, d_r));
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(&h, d_r, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  if (h.mn_enc > h.mx_enc) return 0;  // all NaN (:67-68)
  const float mn = f32_dec(h.mn_enc), mx = f32_dec(h.mx_enc);
  // (:65-66 cannot fire with ignoreZeroVoxels = 0: zeros take part in the minimum)
  if (mn == mx) { *pct2 = mn; *pct98 = mx; return 0; }
  const size_t nZero = (size_t)h.n_nan;  // zeros are not ignored: only NaNs are discounted (:74-76)
  const size_t n2pct = (size_t)round((double)(nvox - nZero) * 0.02);
  if (n2pct < 1 || (nvox - nZero) < 100) { *pct2 = mn; *pct98 = mx; return 0; }
  const int nBins = 1001;
  const float scl = (float)(nBins - 1) / (mx - mn);
  static thread_local unsigned int hist[ISO_MAXBINS];
  B2M_TRY(iso_hist(ctx, d_img, nvox, mn, scl, nBins, 0, hist));
  size_t n = 0, lo = 0;
  while (n < n2pct) { n += hist[lo]; lo++; }
  lo--;
  n = 0;
  int hi = nBins;
  while (n < n2pct) { hi--; n += hist[hi]; }
  if ((int)lo == hi) {  // the majority is neither black nor white (:112-128)
    int ok = -1;
    while (ok != 0) {
      if (lo > 0) { lo--; if (hist[lo] > 0) ok = 0; }
      if (ok != 0 && hi < nBins - 1) { hi++; if (hist[hi] > 0) ok = 0; }
      if (lo == 0 && hi == nBins - 1) ok = 0;
    }
  }
  *pct2 = (float)lo / scl + mn;
  *pct98 = (float)hi / scl + mn;
  return 0;
}

// nii_otsu (isolevel.c:141-243), modes 5 (three thresholds) and 3 (one).  P(u,v) = S(u,v)^2 / P(u,v) of Liao's eq. 29
// is evaluated from the two prefix arrays when needed instead of from a stored nBin x nBin table: the same double
// operations in the same order, hence the same values.
namespace {
struct otsu_tab {
  enum { N = 256 };
  double P[N], S[N];
  double score(int u, int v) const {  // the reference's P[(u*nBin)+v] after the eq.-29 loop, u <= v
    const double p = u == 0 ? P[v] : P[v] - P[u - 1];
    const double s = u == 0 ? S[v] : S[v] - S[u - 1];
    return p != 0 ? (s * s) / p : p;
  }
};
}  // namespace

static void otsu(const unsigned int *H, int mode, int *dark, int *mid, int *bright) {
  const int nBin = otsu_tab::N;
  *dark = *mid = *bright = 0;
  double Sum = 0.0;
  for (int v = 0; v < nBin; v++) Sum = Sum + (int)H[v];
  if (Sum <= 0) return;
  static thread_local otsu_tab T;
  T.P[0] = (int)H[0];  // sic: the first bin enters un-normalised (:159-160)
  T.S[0] = (int)H[0];
  for (int v = 1; v < nBin; v++) {
    const double Prob = (int)H[v] / Sum;
    T.P[v] = T.P[v - 1] + Prob;
    T.S[v] = T.S[v - 1] + (v + 1) * Prob;
  }
  if (mode == 5) {
    // the first-segment and last-segment scores only depend on one index: tabulate them, and the two middle segments
    // per (l, m) / (m, h) pair as the loops run
    static thread_local double first[otsu_tab::N], last[otsu_tab::N];
    for (int i = 0; i < nBin; i++) first[i] = T.score(0, i);
    for (int i = 0; i < nBin - 1; i++) last[i] = T.score(i + 1, 255);
    int lo = (int)(0.25 * nBin), mi = (int)(0.50 * nBin), hi = (int)(0.75 * nBin);
    double max = first[lo] + T.score(lo + 1, mi) + T.score(mi + 1, hi) + last[hi];
    for (int l = 0; l < nBin - 3; l++)
      for (int m = l + 1; m < nBin - 2; m++) {
        const double lm = first[l] + T.score(l + 1, m);
        for (int h = m + 1; h < nBin - 1; h++) {
          const double v = lm + T.score(m + 1, h) + last[h];
          if (v > max) { lo = l; mi = m; hi = h; max = v; }
        }
      }
    *dark = lo; *mid = mi; *bright = hi;
  } else {
    int thresh = (int)(0.25 * nBin);
    double max = T.score(0, thresh) + T.score(thresh + 1, nBin - 1);
    for (int i = 0; i < nBin - 1; i++) {
      const double v = T.score(0, i) + T.score(i + 1, nBin - 1);
      if (v > max) { thresh = i; max = v; }
    }
    *dark = *mid = *bright = thresh;
  }
}

// setThreshold(img, nvox, darkMediumBright123), isolevel.c:245-277, on a device-resident volume
extern "C" int b2m_isolevel_device(b2m_ctx *ctx, const float *d_img, size_t nvox, int dark_medium_bright_123, float *isolevel) {
  if (!ctx || !d_img || !isolevel) { b2m_set_error("null argument"); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  float mn, mx;
  int rr = robust_range(ctx, d_img, nvox, &mn, &mx);
  if (rr < 0) return rr;
  if (rr != 0) { *isolevel = 1.0f; return B2M_OK; }  // "no variability" (:248-249)
  const int kOtsuBins = 256;
  const float scl = (float)(kOtsuBins - 1) / (mx - mn);
  unsigned int hist[256];
  B2M_TRY(iso_hist(ctx, d_img, nvox, mn, scl, kOtsuBins, 1, hist));
  int dark, mid, bright;
  if (dark_medium_bright_123 == 1 || dark_medium_bright_123 == 3) {
    otsu(hist, 5, &dark, &mid, &bright);
    *isolevel = dark_medium_bright_123 == 1 ? ((float)dark / scl) + mn : ((float)bright / scl) + mn;
    return B2M_OK;
  }
  otsu(hist, 3, &dark, &mid, &bright);
  *isolevel = ((float)mid / scl) + mn;
  return B2M_OK;
}

extern "C" int b2m_isolevel_host(b2m_ctx *ctx, const float *h_img, size_t nvox, int dark_medium_bright_123, float *isolevel) {
  if (!ctx || !h_img || !isolevel) { b2m_set_error("null argument"); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  B2M_TRY(b2m_reserve(ctx, BUF_INPUT, nvox * 4 + 4));
  B2M_TRY(b2m_copy_h2d(ctx, ctx->buf[BUF_INPUT].p, h_img, nvox * 4));
  return b2m_isolevel_device(ctx, b2m_ptr<float>(ctx, BUF_INPUT), nvox, dark_medium_bright_123, isolevel);
}
