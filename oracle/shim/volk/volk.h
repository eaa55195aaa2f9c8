/* volk/volk.h -- TEST-INFRASTRUCTURE SHIM, not VOLK.
 *
 * Declares the VOLK kernels the reference's native path calls (fastcard/cardet.c:12-19,
 * fastcard/fastcard.c:91,180, fastdet/corr_detector.cpp:64,132-155, fastdet/fastcard_wrappers.h:56-70);
 * shim_impl.c implements them with the scalar loops of VOLK's "generic" protokernels
 * (sequential float accumulation, first-maximum index).  Only used to build oracle/_ref/.
 */
#ifndef THR_ORACLE_SHIM_VOLK_H
#define THR_ORACLE_SHIM_VOLK_H
#include <errno.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
#include <cmath>
#include <complex>
#include <cstring>
#include <sstream>
typedef std::complex<float> lv_32fc_t;
extern "C" {
#else
#include <complex.h>
typedef float _Complex lv_32fc_t;
#endif
size_t volk_get_alignment(void);
void *volk_malloc(size_t size, size_t alignment);
void volk_free(void *p);
void volk_32f_accumulator_s32f(float *result, const float *input, unsigned int num_points);
void volk_32f_index_max_16u(uint16_t *target, const float *src0, uint32_t num_points);
void volk_32fc_magnitude_squared_32f_a(float *magnitude, const lv_32fc_t *input, unsigned int num_points);
void volk_32fc_conjugate_32fc(lv_32fc_t *out, const lv_32fc_t *in, unsigned int num_points);
void volk_32fc_x2_multiply_32fc(lv_32fc_t *out, const lv_32fc_t *a, const lv_32fc_t *b, unsigned int num_points);
#ifdef __cplusplus
}
#endif
#endif
