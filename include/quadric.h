/* quadric.h - the reference's post-smooth prototype (/root/reference/src/quadric.h:7), implemented over libb2m
 * (b2m_laplacian_hc_host: incidence list + gather kernels on the GPU, bit-identical to src/quadric.c:315-394).
 * nii2mesh calls it after meshify() + apply_sform() for `-s <iterations>` (src/nii2mesh.c:330-335).
 * quadric_simplify_mesh() (sequential edge collapses, src/quadric.c:396-518) is NOT part of this library: keep the
 * reference's own quadric.c for it, compiled with -Dlaplacian_smoothHC=ref_laplacian_smoothHC if both are linked.
 * On failure (no CUDA device - there is no CPU path -, indices out of range) the vertices are left untouched and a
 * message goes to stderr; the reference's function returns void too. */
#ifndef B2M_QUADRIC_H
#define B2M_QUADRIC_H
#include <stdbool.h>
#include "meshtypes.h"
#ifdef __cplusplus
extern "C" {
#endif
void laplacian_smoothHC(vec3d *verts, vec3i *tris, int nvert, int ntri, double alpha, double beta, int iter, bool lockEdges);
#ifdef __cplusplus
}
#endif
#endif
