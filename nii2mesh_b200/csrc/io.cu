// io.cu — the data formats either side of the hot path, on the GPU.
//
//   ingest : the voxel type conversion of load_nii() (/root/reference/src/nii2mesh.c:155-172): u8 / i16 / u16 / f32 raw
//            voxels -> f32, `(raw * scl_slope) + scl_inter` as an f32 product and an f32 sum (no FMA), scl_slope 0 -> 1
//            (:120).  The RAW bytes cross PCIe (1 or 2 bytes per voxel instead of 4) and are widened on the device.
//   egress : apply_sform() (src/meshify.c:1021-1045): every vertex through the three srow vectors in FP64,
//            (x*s0)+(y*s1)+(z*s2)+s3 left to right, and the winding flip (first two indices swapped) when the
//            determinant proxy - the product of the three row sums, rounded to f32 - is negative.  Applied to the
//            device mesh before the D2H copy.
#include "common.cuh"

template <typename T>
__global__ void __launch_bounds__(256) k_ingest(const T *__restrict__ raw, size_t n, float slope, float inter, float *__restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn((float)raw[i], slope), inter);
}

static int ingest_bpp(int datatype) {
  switch (datatype) {
    case 2: return 1;    // DT_UINT8
    case 4: return 2;    // DT_INT16
    case 512: return 2;  // DT_UINT16
    case 16: return 4;   // DT_FLOAT32
    default: return 0;
  }
}

extern "C" int b2m_ingest_device(b2m_ctx *ctx, const void *d_raw, int datatype, size_t nvox, float scl_slope, float scl_inter,
                                 float *d_out) {
  if (!ctx || !d_raw || !d_out) { b2m_set_error("null argument"); return B2M_EARG; }
  if (!ingest_bpp(datatype)) { b2m_set_error("datatype %d: the reference reads u8 (2), i16 (4), u16 (512), f32 (16)", datatype); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  if (scl_slope == 0.0f) scl_slope = 1.0f;  // nii2mesh.c:120
  size_t blocks = (nvox + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 32) blocks = (size_t)ctx->sm_count * 32;
  if (blocks < 1) blocks = 1;
  switch (datatype) {
    case 2: KT_LAUNCH(ctx, "ingest", k_ingest<uint8_t><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint8_t *)d_raw, nvox, scl_slope, scl_inter, d_out)); break;
    case 4: KT_LAUNCH(ctx, "ingest", k_ingest<int16_t><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const int16_t *)d_raw, nvox, scl_slope, scl_inter, d_out)); break;
    case 512: KT_LAUNCH(ctx, "ingest", k_ingest<uint16_t><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const uint16_t *)d_raw, nvox, scl_slope, scl_inter, d_out)); break;
    default: KT_LAUNCH(ctx, "ingest", k_ingest<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const float *)d_raw, nvox, scl_slope, scl_inter, d_out)); break;
  }
  CU_TRY(cudaGetLastError());
  return B2M_OK;
}

extern "C" int b2m_ingest_host(b2m_ctx *ctx, const void *h_raw, int datatype, size_t nvox, float scl_slope, float scl_inter,
                               float *d_out) {
  if (!ctx || !h_raw || !d_out) { b2m_set_error("null argument"); return B2M_EARG; }
  const int bpp = ingest_bpp(datatype);
  if (!bpp) { b2m_set_error("datatype %d: the reference reads u8 (2), i16 (4), u16 (512), f32 (16)", datatype); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  B2M_TRY(b2m_reserve(ctx, BUF_TMP0, nvox * bpp + 16));
  B2M_TRY(b2m_copy_h2d(ctx, ctx->buf[BUF_TMP0].p, h_raw, nvox * bpp));
  B2M_TRY(b2m_ingest_device(ctx, ctx->buf[BUF_TMP0].p, datatype, nvox, scl_slope, scl_inter, d_out));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B2M_OK;
}

struct sform_rows { float x[4], y[4], z[4]; };
__device__ __forceinline__ double sform_row(double px, double py, double pz, const float s[4]) {
  return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, (double)s[0]), __dmul_rn(py, (double)s[1])), __dmul_rn(pz, (double)s[2])), (double)s[3]);
}
__global__ void __launch_bounds__(256) k_sform_verts(double *__restrict__ v, unsigned n, sform_rows r) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = v[3 * (size_t)i], y = v[3 * (size_t)i + 1], z = v[3 * (size_t)i + 2];
  v[3 * (size_t)i] = sform_row(x, y, z, r.x);
  v[3 * (size_t)i + 1] = sform_row(x, y, z, r.y);
  v[3 * (size_t)i + 2] = sform_row(x, y, z, r.z);
}
__global__ void __launch_bounds__(256) k_flip_winding(int *__restrict__ t, unsigned n) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = t[3 * (size_t)i], b = t[3 * (size_t)i + 1];
  t[3 * (size_t)i] = b;
  t[3 * (size_t)i + 1] = a;
}

// in place on the device mesh of `res` (the mesh of the last hot-path call on ctx)
extern "C" int b2m_apply_sform_device(b2m_ctx *ctx, const b2m_result *res, const float srow_x[4], const float srow_y[4],
                                      const float srow_z[4]) {
  if (!ctx || !res || !srow_x || !srow_y || !srow_z) { b2m_set_error("null argument"); return B2M_EARG; }
  CU_TRY(cudaSetDevice(ctx->device));
  sform_rows r;
  for (int k = 0; k < 4; k++) { r.x[k] = srow_x[k]; r.y[k] = srow_y[k]; r.z[k] = srow_z[k]; }
  if (res->nverts > 0)
    KT_LAUNCH(ctx, "sform", k_sform_verts<<<b2m_cdiv((size_t)res->nverts, 256), 256, 0, ctx->stream>>>((double *)res->d_verts, (unsigned)res->nverts, r));
  // determinant proxy exactly as the reference: doubles of float sums, product rounded to f32 (meshify.c:1033-1037)
  const double px = srow_x[0] + srow_x[1] + srow_x[2];
  const double py = srow_y[0] + srow_y[1] + srow_y[2];
  const double pz = srow_z[0] + srow_z[1] + srow_z[2];
  const float det = (float)(px * py * pz);
  if (!(det >= 0.0f) && res->ntris > 0)
    KT_LAUNCH(ctx, "sform", k_flip_winding<<<b2m_cdiv((size_t)res->ntris, 256), 256, 0, ctx->stream>>>((int *)res->d_tris, (unsigned)res->ntris));
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return B2M_OK;
}
