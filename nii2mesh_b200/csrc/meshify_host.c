/* meshify_host.c — plain-C host side of the drop-in boundary: the six functions of the reference's
 * meshify.h (/root/reference/src/meshify.h:7-12), implemented over the libb2m C ABI.
 *
 * meshify() keeps the reference's signature, return values (EXIT_SUCCESS / EXIT_FAILURE), stdout
 * diagnostics and malloc() ownership of *t / *p (src/meshify.c:286-389); the work runs on the GPU
 * through b2m_meshify_host().  It is re-entrant: every host thread gets its own b2m_ctx (the
 * reference is called concurrently from OpenMP threads in atlas mode, src/nii2mesh.c:546-574).
 * apply_sform / strip_ext / clockMsec / timediff / save_mesh are small host utilities.
 */
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/b2m.h"
#include "../../include/meshify.h"
#include "../../include/isolevel.h"
#include "../../include/quadric.h"

int b2m_get_default_backend(void);

static _Thread_local b2m_ctx *tls_ctx = NULL;

static b2m_ctx *get_ctx(void) {
  if (tls_ctx) return tls_ctx;
  int dev = 0;
  const char *e = getenv("B2M_DEVICE");
  if (e) dev = atoi(e);
  if (b2m_create(&tls_ctx, dev) != B2M_OK) {
    fprintf(stderr, "meshify: %s\n", b2m_last_error());
    tls_ctx = NULL;
  }
  return tls_ctx;
}

int meshify(float *img, short dim[3], int originalMC, float isolevel, vec3i **t, vec3d **p, int *nt, int *np,
            bool preSmooth, bool onlyLargest, bool fillBubbles, bool verbose) {
  b2m_ctx *ctx = get_ctx();
  if (!ctx) return EXIT_FAILURE; /* no CUDA device: there is no CPU path */
  int64_t dims[3] = {dim[0], dim[1], dim[2]};
  b2m_opts o;
  memset(&o, 0, sizeof(o));
  o.isolevel = isolevel;
  o.original_mc = originalMC;
  o.pre_smooth = preSmooth;
  o.only_largest = onlyLargest;
  o.fill_bubbles = fillBubbles;
  o.backend = b2m_get_default_backend();
  o.verbose = verbose;
  b2m_result res;
  void *v = NULL, *f = NULL;
  int rc = b2m_meshify_host(ctx, img, dims, &o, &v, &f, &res);
  if (rc != B2M_OK) {
    if (rc < 0) fprintf(stderr, "meshify: %s\n", b2m_last_error());
    return EXIT_FAILURE;
  }
  *t = (vec3i *)f;
  *p = (vec3d *)v;
  *nt = res.ntris;
  *np = res.nverts;
  return EXIT_SUCCESS;
}

/* -i d / m / b (src/isolevel.c:245-277, called from src/nii2mesh.c:586): same prototype, GPU histograms inside */
float setThreshold(float *img, int nvox, int darkMediumBright123) {
  b2m_ctx *ctx = get_ctx();
  float iso = NAN;
  if (!ctx) return iso; /* no CUDA device: there is no CPU path */
  if (nvox < 1) return 1;
  if (b2m_isolevel_host(ctx, img, (size_t)nvox, darkMediumBright123, &iso) != B2M_OK) {
    fprintf(stderr, "setThreshold: %s\n", b2m_last_error());
    return NAN;
  }
  return iso;
}

/* -s <iterations> (src/quadric.c:343-394, called from src/nii2mesh.c:331 and src/obj2mesh.c:122): same prototype, the
 * incidence list and the gathers run on the GPU */
void laplacian_smoothHC(vec3d *verts, vec3i *tris, int nvert, int ntri, double alpha, double beta, int iter, bool lockEdges) {
  b2m_ctx *ctx = get_ctx();
  if (!ctx) return; /* no CUDA device: there is no CPU path */
  if (b2m_laplacian_hc_host(ctx, (double *)verts, (const int *)tris, nvert, ntri, alpha, beta, iter, lockEdges) != B2M_OK)
    fprintf(stderr, "laplacian_smoothHC: %s\n", b2m_last_error());
}

/* voxel -> world transform of the vertices (src/meshify.c:1021-1045).  Each coordinate is the FP64
 * sum (x*s0)+(y*s1)+(z*s2)+s3, left to right; the winding is flipped (first two indices swapped)
 * when the product of the three row sums - the reference's determinant proxy, rounded to f32 -
 * is negative. */
static double affine_row(vec3d v, const float s[4]) { return (v.x * s[0]) + (v.y * s[1]) + (v.z * s[2]) + s[3]; }

void apply_sform(vec3i *t, vec3d *p, int nt, int np, float srow_x[4], float srow_y[4], float srow_z[4]) {
  for (int i = 0; i < np; i++) {
    vec3d v = p[i];
    p[i].x = affine_row(v, srow_x);
    p[i].y = affine_row(v, srow_y);
    p[i].z = affine_row(v, srow_z);
  }
  double sx = srow_x[0] + srow_x[1] + srow_x[2];
  double sy = srow_y[0] + srow_y[1] + srow_y[2];
  double sz = srow_z[0] + srow_z[1] + srow_z[2];
  float det = (float)(sx * sy * sz);
  if (det >= 0.0) return;
  for (int i = 0; i < nt; i++) {
    int a = t[i].x;
    t[i].x = t[i].y;
    t[i].y = a;
  }
}

/* drop the last extension of a path in place: "a/b.nii" -> "a/b", "a/b.nii.gz" -> "a/b.nii"
 * (src/meshify.c:973-982); a leading dot of the file name is not an extension */
void strip_ext(char *fname) {
  char *q = fname + strlen(fname);
  while (q > fname && *q != '.' && *q != '\\' && *q != '/') --q;
  if (q > fname && *q == '.' && q[-1] != '\\' && q[-1] != '/') *q = '\0';
}

double clockMsec(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1000.0 + (ts.tv_nsec / 1.0e6);
}

long timediff(double startTimeMsec, double endTimeMsec) { return (long)round(endTimeMsec - startTimeMsec); }

/* ---- minimal writers: enough to keep the CLI usable (mesh writers are outside the hot path).
 * Build with -DB2M_NO_SAVE_MESH to leave save_mesh to the reference's own writers. ---- */
#ifndef B2M_NO_SAVE_MESH
static int has_ext(const char *fnm, const char *ext) {
  size_t n = strlen(fnm), m = strlen(ext);
  return n >= m && strcmp(fnm + n - m, ext) == 0;
}

static int write_obj(const char *fnm, vec3i *tris, vec3d *pts, int ntri, int npt) {
  FILE *fp = fopen(fnm, "w");
  if (!fp) return EXIT_FAILURE;
  for (int i = 0; i < npt; i++) fprintf(fp, "v %g %g %g\n", pts[i].x, pts[i].y, pts[i].z);
  for (int i = 0; i < ntri; i++) fprintf(fp, "f %d %d %d\n", tris[i].x + 1, tris[i].y + 1, tris[i].z + 1);
  fclose(fp);
  return EXIT_SUCCESS;
}

static int write_ply(const char *fnm, vec3i *tris, vec3d *pts, int ntri, int npt) {
  FILE *fp = fopen(fnm, "wb");
  if (!fp) return EXIT_FAILURE;
  fprintf(fp, "ply\nformat binary_little_endian 1.0\ncomment b2m\nelement vertex %d\n", npt);
  fprintf(fp, "property float x\nproperty float y\nproperty float z\nelement face %d\n", ntri);
  fprintf(fp, "property list uchar int vertex_indices\nend_header\n");
  for (int i = 0; i < npt; i++) {
    float v[3] = {(float)pts[i].x, (float)pts[i].y, (float)pts[i].z};
    fwrite(v, 4, 3, fp);
  }
  for (int i = 0; i < ntri; i++) {
    unsigned char n = 3;
    fwrite(&n, 1, 1, fp);
    fwrite(&tris[i], 4, 3, fp);
  }
  fclose(fp);
  return EXIT_SUCCESS;
}

static int write_mz3(const char *fnm, vec3i *tris, vec3d *pts, int ntri, int npt) {
  /* uncompressed MZ3: 16-byte header (magic 0x5A4D, attr = faces|vertices, nface, nvert, nskip),
   * i32 faces, f32 vertices */
  FILE *fp = fopen(fnm, "wb");
  if (!fp) return EXIT_FAILURE;
  uint16_t magic = 23117, attr = 3;
  uint32_t nf = (uint32_t)ntri, nv = (uint32_t)npt, nskip = 0;
  fwrite(&magic, 2, 1, fp);
  fwrite(&attr, 2, 1, fp);
  fwrite(&nf, 4, 1, fp);
  fwrite(&nv, 4, 1, fp);
  fwrite(&nskip, 4, 1, fp);
  fwrite(tris, 12, (size_t)ntri, fp);
  for (int i = 0; i < npt; i++) {
    float v[3] = {(float)pts[i].x, (float)pts[i].y, (float)pts[i].z};
    fwrite(v, 4, 3, fp);
  }
  fclose(fp);
  return EXIT_SUCCESS;
}

int save_mesh(const char *fnm, vec3i *tris, vec3d *pts, int ntri, int npt, bool isGz) {
  (void)isGz; /* compressed containers are written uncompressed by this shim */
  if (has_ext(fnm, ".obj")) return write_obj(fnm, tris, pts, ntri, npt);
  if (has_ext(fnm, ".ply")) return write_ply(fnm, tris, pts, ntri, npt);
  if (has_ext(fnm, ".mz3")) return write_mz3(fnm, tris, pts, ntri, npt);
  fprintf(stderr, "save_mesh: this build writes .mz3, .ply and .obj only (got %s)\n", fnm);
  return EXIT_FAILURE;
}
#endif /* B2M_NO_SAVE_MESH */
