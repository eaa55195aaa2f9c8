/* fftw3.h -- TEST-INFRASTRUCTURE SHIM, not FFTW.
 *
 * The reference's native path (fastcard/fft.c, fastdet/corr_detector.cpp) links FFTW3f, which is
 * not vendored in /root/reference and not installed here.  This header declares the handful of
 * fftwf_* entry points those sources call (fastcard/fft.c:24-71, fastcard/fastcard.c:32-37,139-144);
 * shim_impl.c implements them with a plain radix-2 FFT evaluated in double precision and rounded to
 * float on output (FFTW's single-precision result is within a few 1e-7 relative of that).
 * It exists only so that oracle/Makefile can compile the reference's own sources into oracle/_ref/.
 */
#ifndef THR_ORACLE_SHIM_FFTW3_H
#define THR_ORACLE_SHIM_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct thr_shim_plan *fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
int fftwf_import_wisdom_from_filename(const char *filename);
int fftwf_export_wisdom_to_filename(const char *filename);
#ifdef __cplusplus
}
#endif
#endif
