/* Minimal stand-in for the legacy TH header, only so that the reference's roi_align.c compiles verbatim
 * into oracle/_ref (we call its pure-C cores ROIAlignForwardCpu directly through ctypes). */
#ifndef OG_TH_SHIM_H
#define OG_TH_SHIM_H
typedef struct { float* data; long size[4]; } THFloatTensor;
static inline float* THFloatTensor_data(THFloatTensor* t) { return t->data; }
static inline long THFloatTensor_size(THFloatTensor* t, int d) { return t->size[d]; }
#endif
