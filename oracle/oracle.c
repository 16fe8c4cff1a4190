/* oracle.c — TEST INFRASTRUCTURE ONLY (never linked, imported or executed by the product path).
 *
 * A plain-C, single-threaded CPU restatement of the voxel->mesh hot path of neurolabusc/nii2mesh
 * (`meshify()`), written from the behaviour of the reference, each function citing the reference
 * file:line it follows (paths relative to /root/reference/).  It is the checker the CUDA kernels
 * are compared against in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * PINNING: this restatement is pinned bit-for-bit against the unmodified reference compiled by
 * oracle/build_ref.sh into oracle/_ref/libref_{lewiner,classic}.so (tests/test_oracle.py)
 * and against golden vectors generated from that build (tests/golden/, tools/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle.c -o liboracle.so -lm
 * (no -march / -ffast-math: FP64/FP32 arithmetic must round exactly like the reference's -O3 build).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MCT_STATIC
#include "../nii2mesh_b200/csrc/mc_tables.inc" /* packed case tables (data only) */

typedef struct { double x, y, z; } ovec3d; /* src/meshtypes.h:4-6 */
typedef struct { int x, y, z; } ovec3i;    /* src/meshtypes.h:8-10 */

/* ------------------------------------------------------------------------------------------
 * Pre-smooth.  src/meshify.c:170-216 (quick_smooth): three separable 5-tap passes x,y,z with
 * double constants 0.05/0.225/0.45, sums left to right in FP64, one rounding to f32 per pass;
 * the two outermost indices on the pass axis keep the previous pass's value.  Returns 1 (and
 * leaves img untouched) when any dim < 5, like the reference's ignored EXIT_FAILURE. */
int orc_smooth(float *img, int nx, int ny, int nz) {
  if (nx < 5 || ny < 5 || nz < 5) return 1;
  size_t n = (size_t)nx * ny * nz;
  float *tmp = (float *)malloc(n * sizeof(float));
  const double w0 = 0.45, w1 = 0.225, w2 = 0.05;
  size_t stride[3] = {1, (size_t)nx, (size_t)nx * ny};
  int len[3] = {nx, ny, nz};
  for (int axis = 0; axis < 3; axis++) {
    size_t s = stride[axis];
    memcpy(tmp, img, n * sizeof(float));
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          int p = axis == 0 ? x : (axis == 1 ? y : z);
          if (p < 2 || p >= len[axis] - 2) continue;
          size_t v = (size_t)x + (size_t)y * nx + (size_t)z * nx * ny;
          double acc = (double)tmp[v - 2 * s] * w2;
          acc = acc + (double)tmp[v - s] * w1;
          acc = acc + (double)tmp[v] * w0;
          acc = acc + (double)tmp[v + s] * w1;
          acc = acc + (double)tmp[v + 2 * s] * w2;
          img[v] = (float)acc;
        }
  }
  free(tmp);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * True connected components with the reference's label numbering.
 * src/bwlabel.c:154-245 (do_initial_labelling: raster scan, backward neighbours),
 * :94-131 (check_previous_slice: face + 4 edge neighbours for conn>=18, + 4 corners for 26),
 * :249-277 (translate_labels: dense labels 1..nl in raster order of each component's first voxel).
 * Unlike the reference's fill_tratab (:54-90) there is no 100-hop give-up: this is the true
 * partition, which is what the reference computes whenever its stderr is clean. */
static uint32_t uf_find(uint32_t *p, uint32_t a) {
  uint32_t r = a;
  while (p[r] != r) r = p[r];
  while (p[a] != r) { uint32_t nx = p[a]; p[a] = r; a = nx; }
  return r;
}
static void uf_union(uint32_t *p, uint32_t a, uint32_t b) {
  a = uf_find(p, a); b = uf_find(p, b);
  if (a == b) return;
  if (a < b) p[b] = a; else p[a] = b;
}
/* bw: N bytes (non-zero = set).  lab: N u32 out (0 = background).  returns number of labels */
int orc_cc_label(const uint8_t *bw, uint32_t *lab, int nx, int ny, int nz, int conn) {
  size_t n = (size_t)nx * ny * nz, nxy = (size_t)nx * ny;
  uint32_t *par = (uint32_t *)malloc((n + 1) * sizeof(uint32_t)); /* provisional label -> parent */
  uint32_t next = 1;
  par[0] = 0;
  memset(lab, 0, n * sizeof(uint32_t));
  /* backward neighbour offsets (dx,dy,dz) */
  static const int8_t nb[13][3] = {
      {-1, 0, 0}, {0, -1, 0}, {0, 0, -1},                                       /* 6  */
      {-1, -1, 0}, {1, -1, 0}, {-1, 0, -1}, {1, 0, -1}, {0, -1, -1}, {0, 1, -1}, /* 18 */
      {-1, -1, -1}, {1, -1, -1}, {-1, 1, -1}, {1, 1, -1}};                      /* 26 */
  int nnb = conn == 6 ? 3 : (conn == 18 ? 9 : 13);
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        size_t v = x + (size_t)y * nx + z * nxy;
        if (!bw[v]) continue;
        uint32_t mine = 0;
        for (int k = 0; k < nnb; k++) {
          int xx = x + nb[k][0], yy = y + nb[k][1], zz = z + nb[k][2];
          if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0) continue;
          uint32_t l = lab[xx + (size_t)yy * nx + zz * nxy];
          if (!l) continue;
          if (!mine) mine = l; else uf_union(par, mine, l);
        }
        if (!mine) { mine = next; par[next] = next; next++; }
        lab[v] = mine;
      }
  /* dense renumber in raster order of first voxel (translate_labels) */
  uint32_t *dense = (uint32_t *)calloc(next + 1, sizeof(uint32_t));
  int nl = 0;
  for (size_t v = 0; v < n; v++) {
    if (!lab[v]) continue;
    uint32_t r = uf_find(par, lab[v]);
    if (!dense[r]) dense[r] = ++nl;
    lab[v] = dense[r];
  }
  free(dense);
  free(par);
  return nl;
}

/* src/bwlabel.c:429-476 (bwlabelCore): label img!=0; onlyLargest keeps the label with the most
 * voxels, strict '>' so ties go to the lowest label (:462-466), output 0/1. */
int orc_bwlabel_core(float *img, int conn, int nx, int ny, int nz, int onlyLargest) {
  if (nx < 2 || ny < 2 || nz < 1) return 0; /* :434-437: refused, img left as it is ("must be 2 or 3-dimensional") */
  size_t n = (size_t)nx * ny * nz;
  uint8_t *bw = (uint8_t *)malloc(n);
  uint32_t *lab = (uint32_t *)malloc(n * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) bw[i] = img[i] != 0.0f;
  int nl = orc_cc_label(bw, lab, nx, ny, nz, conn);
  free(bw);
  if (nl > 0 && onlyLargest) {
    uint32_t *cnt = (uint32_t *)calloc((size_t)nl + 1, sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) cnt[lab[i]]++;
    uint32_t best = 0, bestn = 0;
    for (int j = 1; j <= nl; j++)
      if (cnt[j] > bestn) { bestn = cnt[j]; best = (uint32_t)j; }
    for (size_t i = 0; i < n; i++) lab[i] = lab[i] == best;
    free(cnt);
    nl = 1;
  }
  for (size_t i = 0; i < n; i++) img[i] = (float)lab[i];
  free(lab);
  return nl;
}

/* src/bwlabel.c:478-543 (bwlabel): optional bubble fill = 6-connected labelling of the inverted
 * mask; a background label is "edge" if any of its voxels lies on one of the six volume faces
 * (:498-529); every voxel of a non-edge background label is set (:533-538); skipped entirely when
 * there is at most one background component (:488-491); then the foreground labelling. */
int orc_bwlabel(float *img, int conn, int nx, int ny, int nz, int onlyLargest, int fillBubbles) {
  if (!fillBubbles) return orc_bwlabel_core(img, conn, nx, ny, nz, onlyLargest);
  if (nx < 2 || ny < 2 || nz < 1) return 0; /* both bwlabelCore calls refuse: no bubbles (:488-491), mask untouched */
  size_t n = (size_t)nx * ny * nz, nxy = (size_t)nx * ny;
  uint8_t *bw = (uint8_t *)malloc(n);
  uint32_t *lab = (uint32_t *)malloc(n * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) bw[i] = img[i] == 0.0f;
  int ninv = orc_cc_label(bw, lab, nx, ny, nz, 6);
  free(bw);
  if (ninv > 1) {
    uint8_t *edge = (uint8_t *)calloc((size_t)ninv + 1, 1);
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++)
          if (x == 0 || y == 0 || z == 0 || x == nx - 1 || y == ny - 1 || z == nz - 1)
            edge[lab[x + (size_t)y * nx + z * nxy]] = 1;
    for (size_t i = 0; i < n; i++)
      if (!edge[lab[i]]) img[i] = 1.0f; /* label 0 = foreground: already non-zero */
    free(edge);
  }
  free(lab);
  return orc_bwlabel_core(img, conn, nx, ny, nz, onlyLargest);
}

/* src/meshify.c:218-264 (dilate, is26): interior voxels only; the neighbour loop starts at n=1 so
 * offset (-1,-1,-1) is never tested -> 25-neighbour dilation; reads img, ORs into img afterwards. */
void orc_dilate25(float *img, int nx, int ny, int nz) {
  size_t n = (size_t)nx * ny * nz, nxy = (size_t)nx * ny;
  uint8_t *m = (uint8_t *)calloc(n, 1);
  for (int z = 1; z < nz - 1; z++)
    for (int y = 1; y < ny - 1; y++)
      for (int x = 1; x < nx - 1; x++) {
        size_t v = x + (size_t)y * nx + z * nxy;
        int hit = 0;
        for (int dz = -1; dz <= 1 && !hit; dz++)
          for (int dy = -1; dy <= 1 && !hit; dy++)
            for (int dx = -1; dx <= 1; dx++) {
              if (dx == 0 && dy == 0 && dz == 0) continue;
              if (dx == -1 && dy == -1 && dz == -1) continue;
              if (img[(size_t)((long)v + dx + (long)dy * nx + (long)dz * (long)nxy)] > 0) { hit = 1; break; }
            }
        m[v] = (uint8_t)hit;
      }
  for (size_t v = 1; v < n; v++)
    if (m[v]) img[v] = 1.0f;
  free(m);
}

/* ------------------------------------------------------------------------------------------
 * Front half of meshify(): smooth, min/max, isolevel sanity, CC mask, fill/largest, edge darken,
 * bright bounding box.  src/meshify.c:286-371.  Mutates img exactly like the reference.
 * Returns 0 ok, 1 = "no variability".  iso_io may be reset (:316-319).  mask_out (optional, N
 * floats) receives the bwlabel mask after the optional dilation (debug hook for mask parity). */
int orc_front(float *img, int nx, int ny, int nz, float *iso_io, int preSmooth, int onlyLargest,
              int fillBubbles, int lo[3], int hi[3], float *mn_out, float *mx_out, float *mask_out) {
  size_t n = (size_t)nx * ny * nz;
  float iso = *iso_io;
  if (preSmooth) orc_smooth(img, nx, ny, nz);
  float mx = img[0], mn = img[0];
  for (size_t i = 0; i < n; i++) { mx = fmaxf(mx, img[i]); mn = fminf(mn, img[i]); }
  if (mn_out) *mn_out = mn;
  if (mx_out) *mx_out = mx;
  if (mn == mx) return 1;
  if (iso <= mn || iso > mx) iso = (float)(0.5 * (mn + mx));
  *iso_io = iso;
  if (onlyLargest || fillBubbles) {
    float *mask = (float *)calloc(n, sizeof(float));
    for (size_t i = 0; i < n; i++)
      if (img[i] >= iso) mask[i] = 1.0f;
    orc_bwlabel(mask, 18, nx, ny, nz, onlyLargest, fillBubbles);
    if (fillBubbles)
      for (size_t i = 0; i < n; i++)
        if (mask[i] != 0) img[i] = (float)fmax(img[i], iso);
    if (onlyLargest) {
      orc_dilate25(mask, nx, ny, nz);
      for (size_t i = 0; i < n; i++)
        if (mask[i] == 0) img[i] = mn;
    }
    if (mask_out) memcpy(mask_out, mask, n * sizeof(float));
    free(mask);
  }
  float edgeMax = (float)(0.75 * (mn + iso)); /* f32 add, then double multiply (:346) */
  int dims[3] = {nx, ny, nz};
  for (int a = 0; a < 3; a++) { lo[a] = dims[a]; hi[a] = 0; }
  size_t v = 0;
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++, v++) {
        if (img[v] >= iso) {
          if (x < lo[0]) lo[0] = x;
          if (y < lo[1]) lo[1] = y;
          if (z < lo[2]) lo[2] = z;
          if (x > hi[0]) hi[0] = x;
          if (y > hi[1]) hi[1] = y;
          if (z > hi[2]) hi[2] = z;
        }
        if (x == 0 || y == 0 || z == 0 || x == nx - 1 || y == ny - 1 || z == nz - 1)
          img[v] = fminf(edgeMax, img[v]);
      }
  for (int a = 0; a < 3; a++) {
    lo[a] = lo[a] - 1 > 0 ? lo[a] - 1 : 0;
    hi[a] = hi[a] + 2 < dims[a] ? hi[a] + 2 : dims[a];
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Lewiner MC33.  src/MarchingCubes.c:1086-1143 (wrapper), :235-270 (edge vertices),
 * :116-155 (cube loop), :458-795 (case selection), :276-295 (face test), :301-453 (interior test),
 * :803-859 (triangles), :935-1026 (edge vertex positions), :1029-1082 (centroid vertex).
 * All arithmetic is f32. */
static int mc_face_test(const float *c, int f) { /* :276-295 */
  static const int8_t q[6][4] = {{0, 4, 5, 1}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 4, 0}, {0, 3, 2, 1}, {4, 7, 6, 5}};
  int a = (f < 0 ? -f : f) - 1;
  float A = c[q[a][0]], B = c[q[a][1]], C = c[q[a][2]], D = c[q[a][3]];
  float det = A * C - B * D;
  if (fabsf(det) < FLT_EPSILON) return f >= 0;
  float fa = (float)f * A;
  return fa * det >= 0;
}

static int mc_interior_test(const float *c, int kase, int refedge, int s) { /* :301-453 */
  float t, At = 0, Bt = 0, Ct = 0, Dt = 0;
  if (kase == 4 || kase == 10) { /* :310-322 */
    float a = (c[4] - c[0]) * (c[6] - c[2]) - (c[7] - c[3]) * (c[5] - c[1]);
    float b = c[2] * (c[4] - c[0]) + c[0] * (c[6] - c[2]) - c[1] * (c[7] - c[3]) - c[3] * (c[5] - c[1]);
    t = -b / (2 * a);
    if (t < 0 || t > 1) return s > 0;
    At = c[0] + (c[4] - c[0]) * t;
    Bt = c[3] + (c[7] - c[3]) * t;
    Ct = c[2] + (c[6] - c[2]) * t;
    Dt = c[1] + (c[5] - c[1]) * t;
  } else { /* :324-422: slice at the reference edge's crossing; rows = {e0,e1, B0,B1, C0,C1, D0,D1} */
    static const int8_t e[12][8] = {
        {0, 1, 3, 2, 7, 6, 4, 5}, {1, 2, 0, 3, 4, 7, 5, 6}, {2, 3, 1, 0, 5, 4, 6, 7}, {3, 0, 2, 1, 6, 5, 7, 4},
        {4, 5, 7, 6, 3, 2, 0, 1}, {5, 6, 4, 7, 0, 3, 1, 2}, {6, 7, 5, 4, 1, 0, 2, 3}, {7, 4, 6, 5, 2, 1, 3, 0},
        {0, 4, 3, 7, 2, 6, 1, 5}, {1, 5, 0, 4, 3, 7, 2, 6}, {2, 6, 1, 5, 0, 4, 3, 7}, {3, 7, 2, 6, 1, 5, 0, 4}};
    if (refedge >= 0 && refedge < 12) {
      const int8_t *r = e[refedge];
      t = c[r[0]] / (c[r[0]] - c[r[1]]);
      At = 0;
      Bt = c[r[2]] + (c[r[3]] - c[r[2]]) * t;
      Ct = c[r[4]] + (c[r[5]] - c[r[4]]) * t;
      Dt = c[r[6]] + (c[r[7]] - c[r[6]]) * t;
    }
  }
  int test = (At >= 0) | ((Bt >= 0) << 1) | ((Ct >= 0) << 2) | ((Dt >= 0) << 3);
  switch (test) { /* :428-452 */
    case 5: return (At * Ct - Bt * Dt < FLT_EPSILON) ? s > 0 : s < 0;
    case 10: return (At * Ct - Bt * Dt >= FLT_EPSILON) ? s > 0 : s < 0;
    case 7: case 11: case 13: case 14: case 15: return s < 0;
    default: return s > 0;
  }
}

#define T1(name, cfg) (MCT_BLOB + MCT_##name + (cfg) * MCT_##name##_ROW)
#define T2(name, cfg, sub) (MCT_BLOB + MCT_##name + (cfg) * MCT_##name##_ROW + (sub) * MCT_##name##_SUB)

/* :458-795.  Returns the tiling row, number of triangles, and whether a centroid vertex is needed. */
static const signed char *mc_select(const float *c, int lut, int originalMC, int *ntri, int *needc) {
  *needc = 0;
  if (originalMC) { /* :471-477 */
    const signed char *row = T1(casesClassic, lut);
    int n = 0;
    while (row[3 * n] != -1) n++;
    *ntri = n;
    return row;
  }
  int kase = MCT_BLOB[MCT_cases + 2 * lut], cfg = MCT_BLOB[MCT_cases + 2 * lut + 1];
  int sub = 0;
  switch (kase) {
    case 0: *ntri = 0; return MCT_BLOB;
    case 1: *ntri = 1; return T1(tiling1, cfg);
    case 2: *ntri = 2; return T1(tiling2, cfg);
    case 3:
      if (mc_face_test(c, T1(test3, cfg)[0])) { *ntri = 4; return T1(tiling3_2, cfg); }
      *ntri = 2; return T1(tiling3_1, cfg);
    case 4:
      if (mc_interior_test(c, 4, -1, T1(test4, cfg)[0])) { *ntri = 2; return T1(tiling4_1, cfg); }
      *ntri = 6; return T1(tiling4_2, cfg);
    case 5: *ntri = 3; return T1(tiling5, cfg);
    case 6: {
      const signed char *t = T1(test6, cfg);
      if (mc_face_test(c, t[0])) { *ntri = 5; return T1(tiling6_2, cfg); }
      if (mc_interior_test(c, 6, t[2], t[1])) { *ntri = 3; return T1(tiling6_1_1, cfg); }
      *needc = 1; *ntri = 9; return T1(tiling6_1_2, cfg);
    }
    case 7: {
      const signed char *t = T1(test7, cfg);
      if (mc_face_test(c, t[0])) sub += 1;
      if (mc_face_test(c, t[1])) sub += 2;
      if (mc_face_test(c, t[2])) sub += 4;
      switch (sub) {
        case 0: *ntri = 3; return T1(tiling7_1, cfg);
        case 1: *ntri = 5; return T2(tiling7_2, cfg, 0);
        case 2: *ntri = 5; return T2(tiling7_2, cfg, 1);
        case 3: *needc = 1; *ntri = 9; return T2(tiling7_3, cfg, 0);
        case 4: *ntri = 5; return T2(tiling7_2, cfg, 2);
        case 5: *needc = 1; *ntri = 9; return T2(tiling7_3, cfg, 1);
        case 6: *needc = 1; *ntri = 9; return T2(tiling7_3, cfg, 2);
        default:
          if (mc_interior_test(c, 7, t[4], t[3])) { *ntri = 9; return T1(tiling7_4_2, cfg); }
          *ntri = 5; return T1(tiling7_4_1, cfg);
      }
    }
    case 8: *ntri = 2; return T1(tiling8, cfg);
    case 9: *ntri = 4; return T1(tiling9, cfg);
    case 10: {
      const signed char *t = T1(test10, cfg);
      int f0 = mc_face_test(c, t[0]);
      int f1 = mc_face_test(c, t[1]);
      if (f0 && f1) { *ntri = 4; return T1(tiling10_1_1_, cfg); }
      if (f0) { *needc = 1; *ntri = 8; return T1(tiling10_2, cfg); }
      if (f1) { *needc = 1; *ntri = 8; return T1(tiling10_2_, cfg); }
      if (mc_interior_test(c, 10, -1, t[2])) { *ntri = 4; return T1(tiling10_1_1, cfg); }
      *ntri = 8; return T1(tiling10_1_2, cfg);
    }
    case 11: *ntri = 4; return T1(tiling11, cfg);
    case 12: {
      const signed char *t = T1(test12, cfg);
      int f0 = mc_face_test(c, t[0]);
      int f1 = mc_face_test(c, t[1]);
      if (f0 && f1) { *ntri = 4; return T1(tiling12_1_1_, cfg); }
      if (f0) { *needc = 1; *ntri = 8; return T1(tiling12_2, cfg); }
      if (f1) { *needc = 1; *ntri = 8; return T1(tiling12_2_, cfg); }
      if (mc_interior_test(c, 12, t[3], t[2])) { *ntri = 4; return T1(tiling12_1_1, cfg); }
      *ntri = 8; return T1(tiling12_1_2, cfg);
    }
    case 13: {
      const signed char *t = T1(test13, cfg);
      for (int b = 0; b < 6; b++)
        if (mc_face_test(c, t[b])) sub |= 1 << b;
      int sc = MCT_BLOB[MCT_subconfig13 + sub];
      if (sc == 0) { *ntri = 4; return T1(tiling13_1, cfg); }
      if (sc <= 6) { *ntri = 6; return T2(tiling13_2, cfg, sc - 1); }
      if (sc <= 18) { *needc = 1; *ntri = 10; return T2(tiling13_3, cfg, sc - 7); }
      if (sc <= 22) { *needc = 1; *ntri = 12; return T2(tiling13_4, cfg, sc - 19); }
      if (sc <= 26) {
        int k = sc - 23;
        int refedge = T2(tiling13_5_1, cfg, k)[0];
        if (mc_interior_test(c, 13, refedge, t[6])) { *ntri = 6; return T2(tiling13_5_1, cfg, k); }
        *ntri = 10; return T2(tiling13_5_2, cfg, k);
      }
      if (sc <= 38) { *needc = 1; *ntri = 10; return T2(tiling13_3_, cfg, sc - 27); }
      if (sc <= 44) { *ntri = 6; return T2(tiling13_2_, cfg, sc - 39); }
      if (sc == 45) { *ntri = 4; return T1(tiling13_1_, cfg); }
      *ntri = 0; return MCT_BLOB; /* "Impossible case 13?" (:785) */
    }
    case 14: *ntri = 4; return T1(tiling14, cfg);
  }
  *ntri = 0;
  return MCT_BLOB;
}

typedef struct { float x, y, z; } fv3;

int orc_mc_lewiner(const float *img, int dx, int dy, int dz, const int lo[3], const int hi[3], int originalMC,
                   float isolevel, ovec3d **vs, ovec3i **ts, int *nv, int *nt) {
  int NX = hi[0] - lo[0] + 1, NY = hi[1] - lo[1] + 1, NZ = hi[2] - lo[2] + 1; /* :1088-1090 */
  size_t nin = (size_t)dx * dy * dz, nout = (size_t)NX * NY * NZ, NXY = (size_t)NX * NY;
  float *d = (float *)malloc(nout * sizeof(float));
  float mn = img[0];
  for (size_t i = 0; i < nin; i++) mn = (float)fmin(mn, img[i]);
  mn = mn - isolevel;
  { /* :1101-1115: pad with mn, copy with only a linear-index guard (row wrap quirk) */
    size_t i = 0;
    for (int z = 0; z < NZ; z++)
      for (int y = 0; y < NY; y++) {
        size_t zy = (size_t)(y + lo[1]) * dx + (size_t)(z + lo[2]) * dx * dy;
        for (int x = 0; x < NX; x++, i++) {
          size_t j = (size_t)lo[0] + x + zy;
          d[i] = j < nin ? img[j] - isolevel : mn;
        }
      }
  }
  int *ev[3];
  for (int a = 0; a < 3; a++) {
    ev[a] = (int *)malloc(nout * sizeof(int));
    memset(ev[a], -1, nout * sizeof(int));
  }
  size_t vcap = 65536, tcap = 65536, nvert = 0, ntri = 0;
  fv3 *V = (fv3 *)malloc(vcap * sizeof(fv3));
  ovec3i *T = (ovec3i *)malloc(tcap * sizeof(ovec3i));
#define CLAMPEPS(v) (fabsf(v) < FLT_EPSILON ? FLT_EPSILON : (v))
  /* :235-270 edge vertices, x then y then z per voxel, raster order */
  for (int k = 0; k < NZ; k++)
    for (int j = 0; j < NY; j++)
      for (int i = 0; i < NX; i++) {
        size_t p = i + (size_t)j * NX + k * NXY;
        float c0 = d[p];
        float c[3];
        c[0] = i < NX - 1 ? d[p + 1] : c0;
        c[1] = j < NY - 1 ? d[p + NX] : c0;
        c[2] = k < NZ - 1 ? d[p + NXY] : c0;
        c0 = CLAMPEPS(c0);
        for (int a = 0; a < 3; a++) {
          float c1 = CLAMPEPS(c[a]);
          int cross = c0 < 0 ? c1 > 0 : c1 < 0;
          if (!cross) continue;
          if (nvert >= vcap) { vcap *= 2; V = (fv3 *)realloc(V, vcap * sizeof(fv3)); }
          float den = c0 - c1;
          float u = den != 0.0f ? c0 / den : 0.5f; /* :943-944 */
          fv3 q = {(float)i, (float)j, (float)k};
          if (a == 0) q.x = (float)i + u; else if (a == 1) q.y = (float)j + u; else q.z = (float)k + u;
          ev[a][p] = (int)nvert;
          V[nvert++] = q;
        }
      }
  /* :125-146 cubes */
  for (int k = 0; k < NZ - 1; k++)
    for (int j = 0; j < NY - 1; j++)
      for (int i = 0; i < NX - 1; i++) {
        size_t p = i + (size_t)j * NX + k * NXY;
        float c[8];
        int lut = 0;
        for (int q = 0; q < 8; q++) {
          size_t pp = p + ((q ^ (q >> 1)) & 1) + (size_t)((q >> 1) & 1) * NX + (size_t)((q >> 2) & 1) * NXY;
          c[q] = CLAMPEPS(d[pp]);
          if (c[q] > 0) lut |= 1 << q;
        }
        if (lut == 0 || lut == 255) continue;
        int n3, needc;
        const signed char *row = mc_select(c, lut, originalMC, &n3, &needc);
        /* edge code -> vertex id (:813-825) */
        size_t eo[12] = {p, p + 1, p + NX, p, p + NXY, p + 1 + NXY, p + NX + NXY, p + NXY, p, p + 1, p + 1 + NX, p + NX};
        static const int8_t eax[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};
        int v12 = -1;
        if (needc) { /* :1029-1082 centroid of existing edge vertices in edge-code order, f32 */
          if (nvert >= vcap) { vcap *= 2; V = (fv3 *)realloc(V, vcap * sizeof(fv3)); }
          float u = 0;
          fv3 s = {0, 0, 0};
          for (int e = 0; e < 12; e++) {
            int id = ev[eax[e]][eo[e]];
            if (id == -1) continue;
            ++u;
            s.x += V[id].x; s.y += V[id].y; s.z += V[id].z;
          }
          if (u > 0) { s.x /= u; s.y /= u; s.z /= u; }
          v12 = (int)nvert;
          V[nvert++] = s;
        }
        for (int t = 0; t < n3; t++) {
          int tv[3];
          for (int q = 0; q < 3; q++) {
            int code = row[3 * t + q];
            tv[q] = code == 12 ? v12 : ev[eax[code]][eo[code]];
          }
          if (ntri >= tcap) { tcap *= 2; T = (ovec3i *)realloc(T, tcap * sizeof(ovec3i)); }
          T[ntri].x = tv[2]; T[ntri].y = tv[1]; T[ntri].z = tv[0]; /* winding reversed at export (:1134-1136) */
          ntri++;
        }
      }
  free(d);
  for (int a = 0; a < 3; a++) free(ev[a]);
  if (nvert < 3 || ntri < 1) { free(V); free(T); return 1; }
  *vs = (ovec3d *)malloc(nvert * sizeof(ovec3d));
  for (size_t i = 0; i < nvert; i++) { /* :1127-1129: float + int is an f32 add, then widened */
    (*vs)[i].x = (double)(float)(V[i].x + (float)lo[0]);
    (*vs)[i].y = (double)(float)(V[i].y + (float)lo[1]);
    (*vs)[i].z = (double)(float)(V[i].z + (float)lo[2]);
  }
  free(V);
  *ts = (ovec3i *)realloc(T, ntri * sizeof(ovec3i));
  *nv = (int)nvert;
  *nt = (int)ntri;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Classic marching cubes (triangle soup, FP64).  src/oldcubes.c:465-522 (loop), :50-463
 * (PolygoniseCube: index bit p set iff val[p] < iso :406-414; 12 edge interpolations :428-451;
 * triTable :454-461), :22-40 (VertexInterp: mu=(iso-v1)/(v2-v1); p = p1 + mu*(p2-p1)). */
int orc_mc_classic(const float *img, int dx, int dy, int dz, const int lo[3], const int hi[3], float isolevel,
                   ovec3d **vs, ovec3i **ts, int *nv, int *nt) {
  (void)dz;
  size_t NX = (size_t)dx, NXY = (size_t)dx * dy;
  static const int8_t ea[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, eb[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
  static const int8_t cx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, cy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, cz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
  size_t cap = 1 << 20, np = 0;
  ovec3d *P = (ovec3d *)malloc(cap * sizeof(ovec3d));
  double iso = isolevel;
  for (int z = lo[2]; z < hi[2] - 1; z++)
    for (int y = lo[1]; y < hi[1] - 1; y++)
      for (int x = lo[0]; x < hi[0] - 1; x++) {
        double val[8];
        int idx = 0;
        for (int q = 0; q < 8; q++) {
          val[q] = img[(size_t)(x + cx[q]) + (size_t)(y + cy[q]) * NX + (size_t)(z + cz[q]) * NXY];
          if (val[q] < iso) idx |= 1 << q;
        }
        if (idx == 0 || idx == 255) continue;
        const signed char *row = MCT_BLOB + MCT_casesClassic + 16 * idx;
        if (np + 15 > cap) { cap = cap * 3 / 2; P = (ovec3d *)realloc(P, cap * sizeof(ovec3d)); }
        for (int i = 0; row[i] != -1; i++) {
          int e = row[i], a = ea[e], b = eb[e];
          double mu = (iso - val[a]) / (val[b] - val[a]);
          double ax = (double)x + cx[a], ay = (double)y + cy[a], az = (double)z + cz[a];
          double bx = (double)x + cx[b], by = (double)y + cy[b], bz = (double)z + cz[b];
          P[np].x = ax + mu * (bx - ax);
          P[np].y = ay + mu * (by - ay);
          P[np].z = az + mu * (bz - az);
          np++;
        }
      }
  if (np < 3) { free(P); return 1; }
  *vs = (ovec3d *)realloc(P, np * sizeof(ovec3d));
  int ntri = (int)(np / 3);
  *ts = (ovec3i *)malloc((size_t)ntri * sizeof(ovec3i));
  for (int i = 0; i < ntri; i++) { (*ts)[i].x = 3 * i; (*ts)[i].y = 3 * i + 1; (*ts)[i].z = 3 * i + 2; }
  *nv = (int)np;
  *nt = ntri;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Vertex weld.  src/meshify.c:45-106 (unify_vertices) + src/radixsort.c:277-348 (stable LSD radix
 * sort of (f32 key, index)): key = (float)distance to pts[0]; greedy clustering in key order with
 * tol = 1e-5f: an unassigned i becomes a head; every j>=i with (key_j-key_i) < tol (f32) and FP64
 * distance to the head < tol is (re)assigned; new vertex k takes the coordinates of the highest
 * original index mapped to k; faces remapped.  Returns the new vertex count. */
static double pdist(ovec3d a, ovec3d b) {
  double x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
  return sqrt(x * x + y * y + z * z);
}
int orc_weld(ovec3d **inpt, ovec3i *tris, int npt, int ntri) {
  ovec3d *pts = *inpt;
  float *key = (float *)malloc((size_t)npt * sizeof(float));
  uint32_t *idx = (uint32_t *)malloc((size_t)npt * sizeof(uint32_t));
  uint32_t *idx2 = (uint32_t *)malloc((size_t)npt * sizeof(uint32_t));
  for (int i = 0; i < npt; i++) { key[i] = (float)pdist(pts[0], pts[i]); idx[i] = (uint32_t)i; }
  for (int pass = 0; pass < 4; pass++) { /* stable LSD byte radix on the (non-negative) f32 bit patterns */
    size_t cnt[257] = {0};
    for (int i = 0; i < npt; i++) { uint32_t b; memcpy(&b, &key[idx[i]], 4); cnt[((b >> (8 * pass)) & 255) + 1]++; }
    for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
    for (int i = 0; i < npt; i++) { uint32_t b; memcpy(&b, &key[idx[i]], 4); idx2[cnt[(b >> (8 * pass)) & 255]++] = idx[i]; }
    uint32_t *t = idx; idx = idx2; idx2 = t;
  }
  int *o2n = (int *)malloc((size_t)npt * sizeof(int));
  for (int i = 0; i < npt; i++) o2n[i] = -1;
  const float tol = 0.00001f;
  int nnew = 0;
  for (int i = 0; i < npt; i++) {
    if (o2n[idx[i]] >= 0) continue;
    float k0 = key[idx[i]];
    ovec3d p0 = pts[idx[i]];
    for (int j = i; j < npt && (float)(key[idx[j]] - k0) < tol; j++)
      if (pdist(p0, pts[idx[j]]) < tol) o2n[idx[j]] = nnew;
    nnew++;
  }
  free(key); free(idx); free(idx2);
  if (nnew == npt) { free(o2n); return npt; }
  for (int i = 0; i < ntri; i++) { tris[i].x = o2n[tris[i].x]; tris[i].y = o2n[tris[i].y]; tris[i].z = o2n[tris[i].z]; }
  ovec3d *np_ = (ovec3d *)malloc((size_t)nnew * sizeof(ovec3d));
  for (int i = 0; i < npt; i++) np_[o2n[i]] = pts[i];
  free(pts); free(o2n);
  *inpt = np_;
  return nnew;
}

/* src/meshify.c:113-168 (remove_degenerate_triangles): FP64 side lengths, Kahan's needle-safe Heron
 * formula, drop when c-(a-b) <= 0 or 0.25*sqrt(...) < FLT_EPSILON; order-preserving compaction. */
int orc_degenerate(const ovec3d *pts, ovec3i *tris, int ntri) {
  int k = 0;
  for (int i = 0; i < ntri; i++) {
    double l = pdist(pts[tris[i].x], pts[tris[i].y]);
    double m = pdist(pts[tris[i].x], pts[tris[i].z]);
    double n = pdist(pts[tris[i].y], pts[tris[i].z]);
    double c = fmin(fmin(l, m), n), a = fmax(fmax(l, m), n);
    double b = l + m + n - a - c;
    if ((c - (a - b)) <= 0.0) continue;
    double area4 = 0.25 * sqrt((a + (b + c)) * (c - (a - b)) * (c + (a - b)) * (a + (b - c)));
    if (area4 < FLT_EPSILON) continue;
    tris[k++] = tris[i];
  }
  return k;
}

/* src/meshify.c:286-389 (meshify).  backend: 0 = Lewiner build (originalMC honoured), 1 = classic
 * build (-DUSE_CLASSIC_CUBES; originalMC ignored).  pre_nv/pre_nt receive the pre-weld counts. */
int orc_meshify(float *img, int nx, int ny, int nz, int originalMC, float isolevel, ovec3i **t, ovec3d **p, int *nt,
                int *np, int preSmooth, int onlyLargest, int fillBubbles, int backend, int *pre_nv, int *pre_nt,
                float *iso_used) {
  int lo[3], hi[3];
  float iso = isolevel;
  if (orc_front(img, nx, ny, nz, &iso, preSmooth, onlyLargest, fillBubbles, lo, hi, NULL, NULL, NULL)) return 1;
  if (iso_used) *iso_used = iso;
  ovec3d *pts = NULL;
  ovec3i *tris = NULL;
  int npt = 0, ntri = 0, rc;
  if (backend == 1) rc = orc_mc_classic(img, nx, ny, nz, lo, hi, iso, &pts, &tris, &npt, &ntri);
  else rc = orc_mc_lewiner(img, nx, ny, nz, lo, hi, originalMC, iso, &pts, &tris, &npt, &ntri);
  if (rc) return 1;
  if (pre_nv) *pre_nv = npt;
  if (pre_nt) *pre_nt = ntri;
  npt = orc_weld(&pts, tris, npt, ntri);
  if (npt < 3) return 1;
  ntri = orc_degenerate(pts, tris, ntri);
  *t = tris; *p = pts; *nt = ntri; *np = npt;
  return 0;
}

/* ---- isolevel selection: setThreshold() (/root/reference/src/isolevel.c:245-277) ---------------------------------
 * robust range = 2nd..98th percentile on a 1001-bin histogram (nifti_robust_range :35-139, ignoreZeroVoxels = 0), a
 * 256-bin histogram over it (:251-262), Otsu on that (nii_otsu :141-243: mode 5 -> dark / bright, mode 3 -> medium). */
static int orc_robust_range(const float *img, int nvox, float *pct2, float *pct98) {
  *pct2 = 0.0f; *pct98 = 1.0f;
  if (nvox < 1) return 1;
  float mn = INFINITY, mx = -INFINITY;
  size_t nNan = 0;
  for (int i = 0; i < nvox; i++) {
    if (isnan(img[i])) { nNan++; continue; }
    mn = fminf(img[i], mn);
    mx = fmaxf(img[i], mx);
  }
  if (mn > mx) return 0;
  if (mn == mx) { *pct2 = mn; *pct98 = mx; return 0; }
  size_t nZero = nNan; /* zeros are not ignored (:74-76) */
  size_t n2pct = (size_t)round(((size_t)nvox - nZero) * 0.02);
  if (n2pct < 1 || ((size_t)nvox - nZero) < 100) { *pct2 = mn; *pct98 = mx; return 0; }
  enum { nBins = 1001 };
  float scl = (nBins - 1) / (mx - mn);
  static int hist[nBins];
  for (int i = 0; i < nBins; i++) hist[i] = 0;
  for (int i = 0; i < nvox; i++) {
    if (isnan(img[i])) continue;
    hist[(int)round((img[i] - mn) * scl)]++;
  }
  size_t n = 0, lo = 0;
  while (n < n2pct) { n += hist[lo]; lo++; }
  lo--;
  n = 0;
  int hi = nBins;
  while (n < n2pct) { hi--; n += hist[hi]; }
  if ((int)lo == hi) {
    int ok = -1;
    while (ok != 0) {
      if (lo > 0) { lo--; if (hist[lo] > 0) ok = 0; }
      if (ok != 0 && hi < nBins - 1) { hi++; if (hist[hi] > 0) ok = 0; }
      if (lo == 0 && hi == nBins - 1) ok = 0;
    }
  }
  *pct2 = lo / scl + mn;
  *pct98 = hi / scl + mn;
  return 0;
}

static void orc_otsu(const int *H, int nBin, int mode, int *dark, int *mid, int *bright) {
  *dark = *mid = *bright = 0;
  double Sum = 0.0;
  for (int v = 0; v < nBin; v++) Sum = Sum + H[v];
  if (Sum <= 0) return;
  double *P = (double *)malloc((size_t)nBin * nBin * sizeof(double));
  double *S = (double *)malloc((size_t)nBin * nBin * sizeof(double));
  P[0] = H[0]; S[0] = H[0]; /* sic: un-normalised (:159-160) */
  for (int v = 1; v < nBin; v++) {
    double Prob = H[v] / Sum;
    P[v] = P[v - 1] + Prob;
    S[v] = S[v - 1] + (v + 1) * Prob;
  }
  for (int u = 1; u < nBin; u++)
    for (int v = u; v < nBin; v++) {
      P[u * nBin + v] = P[v] - P[u - 1];
      S[u * nBin + v] = S[v] - S[u - 1];
    }
  for (int u = 0; u < nBin; u++)
    for (int v = u; v < nBin; v++)
      if (P[u * nBin + v] != 0) P[u * nBin + v] = (S[u * nBin + v] * S[u * nBin + v]) / P[u * nBin + v];
  if (mode == 5) {
    int lo = (int)(0.25 * nBin), mi = (int)(0.50 * nBin), hi = (int)(0.75 * nBin);
    double max = P[lo] + P[(lo + 1) * nBin + mi] + P[(mi + 1) * nBin + hi] + P[(hi + 1) * nBin + 255];
    for (int l = 0; l < nBin - 3; l++)
      for (int m = l + 1; m < nBin - 2; m++)
        for (int h = m + 1; h < nBin - 1; h++) {
          double v = P[l] + P[(l + 1) * nBin + m] + P[(m + 1) * nBin + h] + P[(h + 1) * nBin + 255];
          if (v > max) { lo = l; mi = m; hi = h; max = v; }
        }
    *dark = lo; *mid = mi; *bright = hi;
  } else {
    int thresh = (int)(0.25 * nBin);
    double max = P[thresh] + P[(thresh + 1) * nBin + nBin - 1];
    for (int i = 0; i < nBin - 1; i++) {
      double v = P[i] + P[(i + 1) * nBin + nBin - 1];
      if (v > max) { thresh = i; max = v; }
    }
    *dark = *mid = *bright = thresh;
  }
  free(P); free(S);
}

float orc_set_threshold(const float *img, int nvox, int darkMediumBright123) {
  float mn, mx;
  if (orc_robust_range(img, nvox, &mn, &mx) != 0) return 1;
  enum { kOtsuBins = 256 };
  float scl = (kOtsuBins - 1) / (mx - mn);
  int hist[kOtsuBins];
  for (int i = 0; i < kOtsuBins; i++) hist[i] = 0;
  for (int i = 0; i < nvox; i++) {
    if (isnan(img[i])) continue;
    int idx = (int)round((img[i] - mn) * scl);
    idx = idx < kOtsuBins - 1 ? idx : kOtsuBins - 1;
    idx = idx > 0 ? idx : 0;
    hist[idx]++;
  }
  int dark, mid, bright;
  if (darkMediumBright123 == 1 || darkMediumBright123 == 3) {
    orc_otsu(hist, kOtsuBins, 5, &dark, &mid, &bright);
    return darkMediumBright123 == 1 ? (dark / scl) + mn : (bright / scl) + mn;
  }
  orc_otsu(hist, kOtsuBins, 3, &dark, &mid, &bright);
  return (mid / scl) + mn;
}

/* ------------------------------------------------------------------------------------------
 * Post-smooth: Laplacian smoothing with Humphrey's classes.  src/quadric.c:343-394
 * (laplacian_smoothHC) over src/quadric.c:315-341 (laplacian_smooth) and the border rule of
 * src/quadric.c:186-216 (update_mesh, iteration 0).
 * Restated as a GATHER over a vertex -> (triangle, corner) incidence list in triangle order - the
 * order in which the reference's triangle loop adds to sum[v], so every FP64 sum rounds alike:
 *   corner 0 of (p0,p1,p2): sum[p0] += v[p1] + v[p2];  corner 1: v[p0] + v[p2];  corner 2: v[p0] + v[p1]
 *   new v = sum / num (num = 2 per incident corner), vertices without triangles keep their place.
 * HC step (alpha, beta): q = p; p = L(p); b = p - (verts*alpha + q*(1-alpha)); q = L(b);
 *   p = p - (b*beta + q*(1-beta)).  lockEdges: a vertex stays where it was when some vertex's
 * one-ring (all corners of its incident triangles) holds it exactly once. */
typedef struct { int *start, *ref; } orc_inc;  /* ref = tri*4 + corner, ascending per vertex */
static orc_inc orc_incidence(const ovec3i *tris, int nvert, int ntri) {
  orc_inc I;
  I.start = (int *)calloc((size_t)nvert + 1, sizeof(int));
  I.ref = (int *)malloc((size_t)(ntri ? ntri : 1) * 3 * sizeof(int));
  for (int t = 0; t < ntri; t++) { I.start[tris[t].x + 1]++; I.start[tris[t].y + 1]++; I.start[tris[t].z + 1]++; }
  for (int v = 0; v < nvert; v++) I.start[v + 1] += I.start[v];
  int *fill = (int *)malloc((size_t)(nvert ? nvert : 1) * sizeof(int));
  memcpy(fill, I.start, (size_t)nvert * sizeof(int));
  for (int t = 0; t < ntri; t++) {
    const int v[3] = {tris[t].x, tris[t].y, tris[t].z};
    for (int c = 0; c < 3; c++) I.ref[fill[v[c]]++] = t * 4 + c;
  }
  free(fill);
  return I;
}
static void orc_laplacian(const orc_inc *I, const ovec3i *tris, const ovec3d *in, ovec3d *out, int nvert) {
  for (int v = 0; v < nvert; v++) {
    double sx = 0.0, sy = 0.0, sz = 0.0;
    int num = 0;
    for (int k = I->start[v]; k < I->start[v + 1]; k++) {
      const ovec3i *t = &tris[I->ref[k] >> 2];
      const int c = I->ref[k] & 3;
      const int a = c == 0 ? t->y : t->x, b = c == 2 ? t->y : t->z;
      sx = sx + (in[a].x + in[b].x); sy = sy + (in[a].y + in[b].y); sz = sz + (in[a].z + in[b].z);
      num += 2;
    }
    if (num <= 0) { out[v] = in[v]; continue; }
    out[v].x = sx / num; out[v].y = sy / num; out[v].z = sz / num;
  }
}
void orc_laplacian_hc(ovec3d *verts, const ovec3i *tris, int nvert, int ntri, double alpha, double beta, int iter, int lockEdges) {
  const double alpha1 = 1.0 - alpha, beta1 = 1.0 - beta;
  orc_inc I = orc_incidence(tris, nvert, ntri);
  size_t nb = (size_t)(nvert ? nvert : 1) * sizeof(ovec3d);
  ovec3d *p = (ovec3d *)malloc(nb), *q = (ovec3d *)malloc(nb), *b = (ovec3d *)malloc(nb), *lb = (ovec3d *)malloc(nb);
  memcpy(p, verts, (size_t)nvert * sizeof(ovec3d));
  for (int j = 0; j < iter; j++) {
    memcpy(q, p, (size_t)nvert * sizeof(ovec3d));
    orc_laplacian(&I, tris, q, p, nvert);
    for (int i = 0; i < nvert; i++) {
      b[i].x = p[i].x - (verts[i].x * alpha + q[i].x * alpha1);
      b[i].y = p[i].y - (verts[i].y * alpha + q[i].y * alpha1);
      b[i].z = p[i].z - (verts[i].z * alpha + q[i].z * alpha1);
    }
    orc_laplacian(&I, tris, b, lb, nvert);
    for (int i = 0; i < nvert; i++) {
      p[i].x = p[i].x - (b[i].x * beta + lb[i].x * beta1);
      p[i].y = p[i].y - (b[i].y * beta + lb[i].y * beta1);
      p[i].z = p[i].z - (b[i].z * beta + lb[i].z * beta1);
    }
  }
  unsigned char *border = (unsigned char *)calloc((size_t)(nvert ? nvert : 1), 1);
  if (lockEdges) {
    for (int v = 0; v < nvert; v++) {
      const int n = 3 * (I.start[v + 1] - I.start[v]);
      for (int a = 0; a < n; a++) {  /* ids of the one-ring, corner by corner; an id seen exactly once is a border vertex */
        const ovec3i *ta = &tris[I.ref[I.start[v] + a / 3] >> 2];
        const int ida = a % 3 == 0 ? ta->x : (a % 3 == 1 ? ta->y : ta->z);
        int cnt = 0;
        for (int c = 0; c < n; c++) {
          const ovec3i *tc = &tris[I.ref[I.start[v] + c / 3] >> 2];
          cnt += (c % 3 == 0 ? tc->x : (c % 3 == 1 ? tc->y : tc->z)) == ida;
        }
        if (cnt == 1) border[ida] = 1;
      }
    }
  }
  for (int i = 0; i < nvert; i++)
    if (!border[i]) verts[i] = p[i];
  free(border); free(p); free(q); free(b); free(lb); free(I.start); free(I.ref);
}

void orc_free(void *p) { free(p); }
