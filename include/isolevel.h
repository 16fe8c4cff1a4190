/* isolevel.h - the reference's prototype (/root/reference/src/isolevel.h), implemented over libb2m
 * (b2m_isolevel_host: GPU histograms, Otsu search on the host).  Returns 1 when the volume has no voxels, like the
 * reference; NAN when no CUDA device is available (there is no CPU path). */
#ifndef B2M_ISOLEVEL_H
#define B2M_ISOLEVEL_H
#ifdef __cplusplus
extern "C" {
#endif
float setThreshold(float *img, int nvox, int darkMediumBright123);
#ifdef __cplusplus
}
#endif
#endif
