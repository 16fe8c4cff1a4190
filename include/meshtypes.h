/* meshtypes.h — output element types of the drop-in boundary.
 * Same layout as the reference's src/meshtypes.h:4-10 (vec3d = 3 x f64, vec3i = 3 x i32) so that
 * callers of the reference's meshify.h link against this library unchanged. */
#ifndef MESHIFY_TYPES_H
#define MESHIFY_TYPES_H

typedef struct {
	double x, y, z;
} vec3d;

typedef struct {
	int x, y, z;
} vec3i;

#endif
