/* meshify.h — the reference's public API for the voxel->mesh path, served by libb2m.so.
 *
 * These six prototypes are the ones the reference declares at src/meshify.h:7-12 and that
 * nii2() (src/nii2mesh.c:321-356) and main() call; a program written against the reference header
 * links against this library without source changes.
 *
 * Differences in behaviour (documented in INTEGRATION.md):
 *   - meshify() runs on the GPU (device selected by B2M_DEVICE, default 0) and does not modify
 *     `img` (the reference smooths/masks it in place; no caller reads it afterwards,
 *     src/nii2mesh.c:573,589).
 *   - the classic-vs-Lewiner marching cubes choice, a compile-time switch in the reference
 *     (-DUSE_CLASSIC_CUBES, src/meshify.c:25-29), is a runtime switch: environment variable
 *     B2M_CLASSIC_CUBES=1 or b2m_set_default_backend() (include/b2m.h).
 *   - *t and *p are malloc() blocks owned by the caller, exactly as in the reference
 *     (src/nii2mesh.c:353-354; quadric_simplify_mesh free()s them, src/quadric.c:402,412).
 *   - ARRAY ORDER: triangles come in the reference's order (cube raster order, the removed ones dropped).  Vertices
 *     come in the reference's marching-cubes EMISSION order with the welded-away ones dropped, whereas the reference's
 *     unify_vertices() renumbers the vertices by its radix-sort key (distance from the first vertex) whenever at least
 *     one pair merges (src/meshify.c:88-100).  The mesh is the same - same positions, same triangles after relabelling:
 *     that is what the tests compare, after canonical sorting - but the arrays are not element-for-element identical, so
 *     order-dependent consumers (the reference's quadric_simplify_mesh(), src/quadric.c:396-518, run for -r < 1) can
 *     produce a different, equally valid simplification.  Sorting 43 M vertices by key and relabelling 87 M triangles
 *     would cost about a third of the whole 1024^3 step, for an order no caller of meshify() depends on.
 */
#ifndef MESHIFY_H
#define MESHIFY_H

#include <stdbool.h>
#include "meshtypes.h"

#ifdef __cplusplus
extern "C" {
#endif

void strip_ext(char *fname);
int save_mesh(const char *fnm, vec3i *tris, vec3d *pts, int ntri, int npt, bool isGz);
int meshify(float *img, short dim[3], int originalMC, float isolevel, vec3i **t, vec3d **p, int *nt, int *np,
            bool preSmooth, bool onlyLargest, bool fillBubbles, bool verbose);
void apply_sform(vec3i *t, vec3d *p, int nt, int np, float srow_x[4], float srow_y[4], float srow_z[4]);
double clockMsec(void);
long timediff(double startTimeMsec, double endTimeMsec);

#ifdef __cplusplus
}
#endif

#endif /* MESHIFY_H */
