/* shim_impl.c -- TEST INFRASTRUCTURE: stand-ins for the FFTW3f / VOLK / librtlsdr entry points the
 * reference's native sources call, so that those sources compile here (see fftw3.h, volk/volk.h). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fftw3.h"

struct thr_shim_plan {
    int n, sign;
    float *in, *out;      /* interleaved re,im */
    double *wr, *wi;      /* n/2 twiddles */
    double *buf;          /* 2n work */
    int *rev;
};

void *fftwf_malloc(size_t n) {
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 64) != 0) return NULL;
    return p;
}
void fftwf_free(void *p) { free(p); }
int fftwf_import_wisdom_from_filename(const char *f) { (void)f; return 1; }
int fftwf_export_wisdom_to_filename(const char *f) { (void)f; return 1; }

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags) {
    (void)flags;
    if (n < 1 || (n & (n - 1)) != 0) return NULL;           /* powers of two only */
    struct thr_shim_plan *p = calloc(1, sizeof *p);
    if (!p) return NULL;
    p->n = n;
    p->sign = sign;
    p->in = (float *)in;
    p->out = (float *)out;
    p->wr = malloc(sizeof(double) * (n / 2 + 1));
    p->wi = malloc(sizeof(double) * (n / 2 + 1));
    p->buf = malloc(sizeof(double) * 2 * n);
    p->rev = malloc(sizeof(int) * n);
    int bits = 0;
    while ((1 << bits) < n) ++bits;
    for (int i = 0; i < n; ++i) {
        int r = 0;
        for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
        p->rev[i] = r;
    }
    for (int k = 0; k < n / 2; ++k) {
        p->wr[k] = cos(2.0 * M_PI * k / n);
        p->wi[k] = (sign < 0 ? -1.0 : 1.0) * sin(2.0 * M_PI * k / n);
    }
    return p;
}

void fftwf_execute(const fftwf_plan p) {
    const int n = p->n;
    double *a = p->buf;
    for (int i = 0; i < n; ++i) {
        a[2 * p->rev[i]] = p->in[2 * i];
        a[2 * p->rev[i] + 1] = p->in[2 * i + 1];
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; ++k) {
                const double wr = p->wr[k * step], wi = p->wi[k * step];
                double *u = a + 2 * (i + k), *v = a + 2 * (i + k + half);
                const double tr = v[0] * wr - v[1] * wi, ti = v[0] * wi + v[1] * wr;
                v[0] = u[0] - tr;
                v[1] = u[1] - ti;
                u[0] += tr;
                u[1] += ti;
            }
    }
    for (int i = 0; i < 2 * n; ++i) p->out[i] = (float)a[i];
}

void fftwf_destroy_plan(fftwf_plan p) {
    if (!p) return;
    free(p->wr);
    free(p->wi);
    free(p->buf);
    free(p->rev);
    free(p);
}

/* ---- VOLK generic protokernels */
size_t volk_get_alignment(void) { return 64; }
void *volk_malloc(size_t size, size_t alignment) {
    void *p = NULL;
    if (posix_memalign(&p, alignment < sizeof(void *) ? sizeof(void *) : alignment, size ? size : alignment) != 0)
        return NULL;
    return p;
}
void volk_free(void *p) { free(p); }

void volk_32f_accumulator_s32f(float *result, const float *input, unsigned int num_points) {
    float acc = 0.f;
    for (unsigned int i = 0; i < num_points; ++i) acc += input[i];
    *result = acc;
}
void volk_32f_index_max_16u(uint16_t *target, const float *src0, uint32_t num_points) {
    if (num_points > 65535u) num_points = 65535u;            /* VOLK clamps to USHRT_MAX */
    if (num_points == 0) return;
    float max = src0[0];
    uint16_t index = 0;
    for (uint32_t i = 1; i < num_points; ++i)
        if (src0[i] > max) {
            index = (uint16_t)i;
            max = src0[i];
        }
    *target = index;
}
void volk_32fc_magnitude_squared_32f_a(float *magnitude, const float *input, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; ++i) {
        const float re = input[2 * i], im = input[2 * i + 1];
        magnitude[i] = re * re + im * im;
    }
}
void volk_32fc_conjugate_32fc(float *out, const float *in, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; ++i) {
        out[2 * i] = in[2 * i];
        out[2 * i + 1] = -in[2 * i + 1];
    }
}
void volk_32fc_x2_multiply_32fc(float *out, const float *a, const float *b, unsigned int num_points) {
    for (unsigned int i = 0; i < num_points; ++i) {
        const float ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
        out[2 * i] = ar * br - ai * bi;
        out[2 * i + 1] = ar * bi + ai * br;
    }
}

