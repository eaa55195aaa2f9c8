// thrifty_b200.cu -- C ABI (include/thrifty_b200.h) over the fused detect kernel.
//
// Host-side duties: validate the DetectorSettings (same checks as the reference:
// carrier_detect.py:47-49 window range, soa_estimator.py:33 history >= template-1),
// build conj(FFT(template || 0))/N in float64 (soa_estimator.py:63-76) and lay it out in
// the kernel's digit-reversed order, own device scratch/staging, launch.
//
// The product path has no CPU fallback: every entry point fails with THR_ERR_* if the
// device or the kernel is unavailable.

#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <charconv>
#include <cmath>
#include <complex>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>            // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost ~nothing otherwise

#include "../../include/thrifty_b200.h"
#include "detect_kernel.cuh"
#include "variants.h"
#include "card_ingest.cuh"

using thr::DetectParams;
using thr::Variant;

namespace {

// NVTX range for one phase of a host-buffer call (stage / H2D / launch / D2H + records), closed at scope exit
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_create_error;

struct Slot {                 // one in-flight chunk of the host-buffer API
    cudaStream_t stream = nullptr;
    uint8_t *d_in = nullptr;  // raw u8 staging (max_batch * 2N)
    int64_t *d_idx = nullptr;
    thr_record *d_out = nullptr;
    float *d_iq = nullptr;    // complex64 staging (lazily, c64_chunk * N * 8)
    uint8_t *d_text = nullptr;   // .card text staging (lazily, host_chunk lines)
    int64_t *d_off = nullptr;    // payload offsets inside d_text
    size_t text_cap = 0;
    // Records come back through page-locked staging: a device-to-host copy into the caller's (usually pageable)
    // buffer would block the host until the whole chunk has run, and the next chunk's input copy could not be
    // queued behind it (the .card path ran at 35 instead of 50 GB/s of text for that reason).
    thr_record *h_out = nullptr;     // page-locked, max_batch * n_templates records
    int64_t *h_idx = nullptr;        // page-locked, 2 * max_batch (block indices, payload offsets)
    thr_record *pending_dst = nullptr;
    size_t pending_n = 0;
    uint8_t *h_in = nullptr;         // page-locked input staging for pageable callers (lazily, h_in_cap bytes)
    size_t h_in_cap = 0;
};

}  // namespace

struct thr_detector {
    thr_config cfg;
    Variant var;
    Variant var_generic;                 // generic kernel of the same block length (debug launches) when var is a
    bool has_generic = false;            //   specialised one (detect2x_kernel)
    int device = 0;
    int sm_count = 0;
    int ctas_per_sm = 1;
    int grid = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;       // stream used by the *_device entry points
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float2 *d_tpl = nullptr;
    float2 *d_tpl_generic = nullptr;     // template spectrum in the generic kernel's order (debug launches of a 2x detector)
    float2 *d_tpl_shift = nullptr;       // fastdet semantics: per-carrier-bin shifted template spectra
    float2 *d_tpl_nat = nullptr;         // fastdet semantics: template spectrum in natural order
    float *d_tpl_energy = nullptr;
    float2 *d_scratch = nullptr;
    float2 *d_xsave = nullptr;
    unsigned int *d_bad = nullptr;       // invalid base64 character counter (.card ingest)
    // stage-boundary entry points (thr_sync_batch / thr_soa_batch): kernels and staging, set up on first use
    Variant var_stage[3];                // [1] = stop after FFT#2, [2] = start at the shifted spectrum
    int grid_stage[3] = {0, 0, 0};
    float2 *d_stage_fft = nullptr;       // [stage_chunk][N] shifted spectra
    float2 *d_stage_corr = nullptr;      // [stage_chunk][corr_len] correlations
    int stage_chunk = 0;
    Slot slot[2];
    int c64_chunk = 0;
    int host_chunk = 0;                  // blocks per pipelined chunk of the host-buffer API
    DetectParams base;                   // constant part of the kernel parameters
    int64_t launches = 0;
    std::string err;
    char device_name[64];
};

namespace {

int fail(thr_detector *d, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (d) d->err = buf; else g_create_error = buf;
    return code;
}

#define CU(d, call)                                                                          \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail((d), THR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));  \
    } while (0)

// float64 radix-2 FFT (host, setup only)
void fft_f64(std::vector<std::complex<double>> &a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const double pi = 3.14159265358979323846;
    for (size_t len = 2; len <= n; len <<= 1) {
        std::vector<std::complex<double>> w(len / 2);
        for (size_t k = 0; k < len / 2; ++k)
            w[k] = std::complex<double>(std::cos(2 * pi * k / len), -std::sin(2 * pi * k / len));
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w[k];
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

// carrier_detect.py:17-58 fft_range_index
bool range_index(int start, int stop, int length, int *s, int *e) {
    if (std::abs(start) >= length || std::abs(stop) >= length) return false;
    if (start < 0 && stop >= 0) { start += length; stop += length; }
    if (start < 0) start += length;
    if (stop < 0) stop += length;
    if (stop < start) std::swap(start, stop);
    *s = start;
    *e = stop;
    return true;
}

int launch(thr_detector *d, cudaStream_t st, const uint8_t *d_raw, const float *d_iq, const int64_t *d_idx,
           int n_blocks, thr_record *d_out, float2 *dbg_sfft, float2 *dbg_corr, float *dbg_mag,
           bool allow_overlap = false, int64_t raw_stride = 0) {
    if (n_blocks <= 0) return THR_OK;
    DetectParams p = d->base;
    const int grid = n_blocks < d->grid ? n_blocks : d->grid;
    // Kernels with per-CTA global scratch (global-memory FFT buffer, the 2 x 16384 kernel's parking area, the
    // multi-template X' save area) index it by blockIdx.x, and two launches can be on the device at once: the two
    // pipelined slots of a host-buffer call run on two streams, and with programmatic dependent launch the next batch's
    // CTA k starts while this batch's CTA k may still run.  Two scratch sets: one per slot, or alternating per launch --
    // the latter only for full grids, where launch k+2 cannot start before launch k has left the SMs.
    bool pdl = allow_overlap && (d->cfg.flags & THR_CFG_OVERLAP_LAUNCHES);
    int set = (st == d->slot[1].stream) ? 1 : 0;
    if (pdl && (p.scratch || p.xsave)) {
        if (grid == d->grid) set = (int)(d->launches & 1);
        else pdl = false;
    }
    if (set) {
        const size_t stride = (size_t)d->grid * d->cfg.block_len;
        if (p.scratch) p.scratch += stride;
        if (p.xsave) p.xsave += stride;
    }
    p.raw = d_raw;
    p.raw_stride = raw_stride ? raw_stride : 2 * (int64_t)d->cfg.block_len;
    p.iq = reinterpret_cast<const float2 *>(d_iq);
    p.block_idx = d_idx;
    p.out = d_out;
    p.n_blocks = n_blocks;
    p.dbg_shifted_fft = dbg_sfft;
    p.dbg_corr = dbg_corr;
    p.dbg_fft_mag = dbg_mag;
    const Variant *var = &d->var;
    if (d->has_generic && (dbg_sfft || dbg_corr || dbg_mag)) {    // intermediates: generic kernel (one block)
        var = &d->var_generic;
        p.tpl_spec = d->d_tpl_generic;
    }
    void *args[] = {&p};
    cudaLaunchConfig_t lc;
    std::memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(var->launch_threads);       // workers (+ service warpgroup), see detect_kernel.cuh
    lc.dynamicSmemBytes = var->smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = attr;
        lc.numAttrs = 1;
    }
    CU(d, cudaLaunchKernelExC(&lc, var->fn, args));
    d->launches++;
    return THR_OK;
}

}  // namespace

extern "C" {

int thr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *thr_last_error(const thr_detector *det) {
    return det ? det->err.c_str() : g_create_error.c_str();
}

void thr_destroy(thr_detector *d) {
    if (!d) return;
    cudaSetDevice(d->device);
    for (auto &s : d->slot) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.d_in);
        cudaFree(s.d_idx);
        cudaFree(s.d_out);
        cudaFree(s.d_iq);
        cudaFree(s.d_text);
        cudaFree(s.d_off);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.h_idx) cudaFreeHost(s.h_idx);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (d->own_stream) { cudaStreamSynchronize(d->own_stream); cudaStreamDestroy(d->own_stream); }
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    cudaFree(d->d_tpl);
    cudaFree(d->d_tpl_shift);
    cudaFree(d->d_tpl_generic);
    cudaFree(d->d_tpl_nat);
    cudaFree(d->d_tpl_energy);
    cudaFree(d->d_scratch);
    cudaFree(d->d_xsave);
    cudaFree(d->d_bad);
    cudaFree(d->d_stage_fft);
    cudaFree(d->d_stage_corr);
    delete d;
}

int thr_create(const thr_config *cfg, thr_detector **out) {
    if (!cfg || !out) return fail(nullptr, THR_ERR_INVALID, "null argument");
    *out = nullptr;
    const int N = cfg->block_len, H = cfg->history_len, L = cfg->template_len, NT = cfg->n_templates;
    Variant var;
    if (!thr::pick_variant_single(N, &var))
        return fail(nullptr, THR_ERR_INVALID, "unsupported block_len %d (power of two 1024..32768)", N);
    if (NT < 1 || NT > 32) return fail(nullptr, THR_ERR_INVALID, "n_templates must be 1..32");
    if (NT > 1 && !thr::pick_variant_multi(N, &var))
        return fail(nullptr, THR_ERR_INVALID, "no multi-template kernel for block_len %d in this build", N);
    if (!cfg->templates || L < 1 || L > N) return fail(nullptr, THR_ERR_INVALID, "bad template (len %d)", L);
    if (H < L - 1 || H >= N)   // soa_estimator.py:33 assert history_len >= template_len - 1
        return fail(nullptr, THR_ERR_INVALID, "history_len %d must satisfy template_len-1 <= H < block_len", H);
    if (cfg->carrier_len < 1) return fail(nullptr, THR_ERR_INVALID, "carrier_len must be positive");
    if (cfg->max_batch < 1) return fail(nullptr, THR_ERR_INVALID, "max_batch must be positive");
    if (cfg->flags & ~(THR_CFG_OVERLAP_LAUNCHES | THR_CFG_FASTDET_SEMANTICS | THR_CFG_GENERIC_KERNEL))
        return fail(nullptr, THR_ERR_INVALID, "unknown flags 0x%x", cfg->flags);
    const bool fastdet = (cfg->flags & THR_CFG_FASTDET_SEMANTICS) != 0;
    if (fastdet) {
        if (NT != 1) return fail(nullptr, THR_ERR_INVALID, "fastdet semantics take exactly one template");
        if (cfg->carrier_thresh[2] != 0.0 || cfg->corr_thresh[2] != 0.0)
            return fail(nullptr, THR_ERR_INVALID, "fastdet semantics have no stddev threshold term");
        if (cfg->window_start < 0 && cfg->window_stop >= 0)   // fastcard/cardet.c:44-48
            return fail(nullptr, THR_ERR_INVALID, "Carrier frequency window range not supported.");
        if (!thr::pick_variant_fastdet(N, &var))
            return fail(nullptr, THR_ERR_INVALID, "no fastdet-semantics kernel for block_len %d in this build", N);
    }
    int ws, we;
    if (!range_index(cfg->window_start, cfg->window_stop, N, &ws, &we))   // carrier_detect.py:47-49
        return fail(nullptr, THR_ERR_INVALID, "Frequency window out of range: %d - %d", cfg->window_start,
                    cfg->window_stop);

    // block_len 32768: two interleaved 16384-point transforms in shared memory (detect_kernel_2x.cuh; FFT#1 pruned or in
    // full, one or several templates) instead of the generic global-scratch variant, which stays for debug launches
    // pruned FFT#1: the carrier window and its +-3 fit neighbours span at most 128 consecutive bins (mod N), no stddev
    // term.  The band starts at bin 0 when the window lies in [3,124] (no pre-shift), else 3 bins below the window.
    const int wlen_cfg = (we - ws + 1) > N ? N : (we - ws + 1);
    const bool zoom_cfg = (!fastdet && wlen_cfg + 6 <= 128 && cfg->carrier_thresh[2] == 0.0 && N >= 4096);
    const int zoom_base = (ws >= 3 && we + 3 < 128) ? 0 : ((ws % N) - 3 + N) % N;
    Variant var_generic = var;
    bool use_2x = false;
    if (!fastdet && N == 32768 && !(cfg->flags & THR_CFG_GENERIC_KERNEL)) {
        Variant v2;
        if (thr::pick_variant_2x(N, NT > 1, &v2)) {
            var = v2;
            use_2x = true;
        }
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, THR_ERR_NO_DEVICE, "no CUDA device available");
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, THR_ERR_INVALID, "device ordinal %d out of range (%d devices)", cfg->device, ndev);

    thr_detector *d = new thr_detector();
    d->cfg = *cfg;
    d->cfg.templates = nullptr;
    d->var = var;
    d->var_generic = var_generic;
    d->has_generic = use_2x;
    d->device = cfg->device;
    auto bail = [&](int code) { g_create_error = d->err; thr_destroy(d); return code; };
#define CUC(call)                                                                            \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            fail(d, THR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));           \
            return bail(THR_ERR_CUDA);                                                       \
        }                                                                                    \
    } while (0)

    CUC(cudaSetDevice(d->device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, d->device));
    std::snprintf(d->device_name, sizeof d->device_name, "%.63s", prop.name);
    if (prop.major != 10) {
        fail(d, THR_ERR_NO_DEVICE, "device %d (%s, sm_%d%d) is not sm_100: this library only carries sm_100a code",
             d->device, prop.name, prop.major, prop.minor);
        return bail(THR_ERR_NO_DEVICE);
    }
    d->sm_count = prop.multiProcessorCount;
    CUC(cudaFuncSetAttribute(var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)var.smem));
    int occ = 0;
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, var.fn, var.launch_threads, var.smem));
    if (occ < 1) {
        fail(d, THR_ERR_CUDA, "kernel %s does not fit on an SM (smem %zu)", var.name, var.smem);
        return bail(THR_ERR_CUDA);
    }
    // setmaxnreg only moves registers inside the pool the CTA was launched with; a kernel whose targets exceed it does
    // not fail, it dead-locks -- refuse such a build here (detect_kernel.cuh asserts the same at compile time, assuming
    // ptxas allocates what __launch_bounds__ allows)
    for (const Variant *v : {&var, &var_generic}) {
        if (!v->worker_regs) continue;
        cudaFuncAttributes fa;
        CUC(cudaFuncGetAttributes(&fa, v->fn));
        const int workers = v->launch_threads - 128;
        if (fa.numRegs * v->launch_threads < workers * v->worker_regs + 128 * 32) {
            fail(d, THR_ERR_CUDA, "kernel %s was built with %d registers x %d threads, fewer than its setmaxnreg split needs "
                 "(%d x %d + 128 x 32)", v->name, fa.numRegs, v->launch_threads, workers, v->worker_regs);
            return bail(THR_ERR_CUDA);
        }
    }
    d->ctas_per_sm = occ;
    d->grid = d->sm_count * occ;
    if (const char *cap = std::getenv("THRIFTY_B200_MAX_GRID")) {      // testing aid: few CTAs walk many blocks each
        const int g = std::atoi(cap);                                  // (compute-sanitizer runs on small inputs)
        if (g >= 1 && g < d->grid) d->grid = g;
    }
    if (use_2x)
        CUC(cudaFuncSetAttribute(var_generic.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)var_generic.smem));

    CUC(cudaStreamCreateWithFlags(&d->own_stream, cudaStreamNonBlocking));
    d->stream = d->own_stream;
    CUC(cudaEventCreate(&d->ev0));
    CUC(cudaEventCreate(&d->ev1));

    // ---- template spectra: conj(FFT(template || zeros))/N, float64 -> float32, kernel order
    {
        const int T = var.threads, R2 = var.r2, R3 = var.r3, I3 = var.i3;
        std::vector<float2> perm((size_t)NT * N), perm_generic;
        std::vector<float> energy(NT);
        std::vector<std::complex<double>> a(N);
        for (int t = 0; t < NT; ++t) {
            double e = 0.0;
            for (int i = 0; i < N; ++i) {
                const double v = i < L ? cfg->templates[(size_t)t * L + i] : 0.0;
                a[i] = v;
                e += v * v;
            }
            energy[t] = (float)e;
            fft_f64(a);
            for (int i = 0; i < I3; ++i)
                for (int k3 = 0; k3 < R3; ++k3)
                    for (int tid = 0; tid < T; ++tid) {
                        const int g = var.p3_item(tid, i);
                        const int k = (g / R2) + 32 * (g % R2) + 32 * R2 * k3;
                        const size_t pos = (size_t)t * N + (size_t)(i * R3 + k3) * T + tid;
                        perm[pos] = make_float2((float)(a[k].real() / N), (float)(-a[k].imag() / N));
                        if (var.two_halves)      // bins k >= N/2 in the same thread order, after the first half
                            perm[pos + N / 2] = make_float2((float)(a[k + N / 2].real() / N), (float)(-a[k + N / 2].imag() / N));
                    }
            if (use_2x) {                        // the generic kernel's own order for debug launches
                const int Tg = var_generic.threads, R2g = var_generic.r2, R3g = var_generic.r3, I3g = var_generic.i3;
                perm_generic.resize((size_t)NT * N);
                for (int i = 0; i < I3g; ++i)
                    for (int k3 = 0; k3 < R3g; ++k3)
                        for (int tid = 0; tid < Tg; ++tid) {
                            const int g = var_generic.p3_item(tid, i);
                            const int k = (g / R2g) + 32 * (g % R2g) + 32 * R2g * k3;
                            perm_generic[(size_t)t * N + (size_t)(i * R3g + k3) * Tg + tid] =
                                make_float2((float)(a[k].real() / N), (float)(-a[k].imag() / N));
                        }
            }
        }
        if (fastdet) {
            // a is still FFT(template 0): natural-order conj/N, and one pre-rolled copy per carrier bin of the
            // window in kernel order (fastdet/corr_detector.cpp:13-17,179: roll(fft, -argmax) == re-index T)
            const int wlen = (we - ws + 1) > N ? N : (we - ws + 1);
            std::vector<float2> nat(N);
            for (int k = 0; k < N; ++k) nat[k] = make_float2((float)(a[k].real() / N), (float)(-a[k].imag() / N));
            CUC(cudaMalloc(&d->d_tpl_nat, nat.size() * sizeof(float2)));
            CUC(cudaMemcpy(d->d_tpl_nat, nat.data(), nat.size() * sizeof(float2), cudaMemcpyHostToDevice));
            if ((size_t)wlen * N * sizeof(float2) <= ((size_t)256 << 20)) {
                std::vector<float2> sh((size_t)wlen * N);
                for (int r = 0; r < wlen; ++r) {
                    const int kpeak = (ws + r) % N;
                    for (int i = 0; i < I3; ++i)
                        for (int k3 = 0; k3 < R3; ++k3)
                            for (int tid = 0; tid < T; ++tid) {
                                const int g = var.p3_item(tid, i);
                                const int k = (g / R2) + 32 * (g % R2) + 32 * R2 * k3;
                                sh[(size_t)r * N + (size_t)(i * R3 + k3) * T + tid] = nat[(k - kpeak + N) % N];
                            }
                }
                CUC(cudaMalloc(&d->d_tpl_shift, sh.size() * sizeof(float2)));
                CUC(cudaMemcpy(d->d_tpl_shift, sh.data(), sh.size() * sizeof(float2), cudaMemcpyHostToDevice));
            }
        }
        CUC(cudaMalloc(&d->d_tpl, perm.size() * sizeof(float2)));
        CUC(cudaMemcpy(d->d_tpl, perm.data(), perm.size() * sizeof(float2), cudaMemcpyHostToDevice));
        if (use_2x) {
            CUC(cudaMalloc(&d->d_tpl_generic, perm_generic.size() * sizeof(float2)));
            CUC(cudaMemcpy(d->d_tpl_generic, perm_generic.data(), perm_generic.size() * sizeof(float2), cudaMemcpyHostToDevice));
        }
        CUC(cudaMalloc(&d->d_tpl_energy, NT * sizeof(float)));
        CUC(cudaMemcpy(d->d_tpl_energy, energy.data(), NT * sizeof(float), cudaMemcpyHostToDevice));
    }
    // per-CTA scratch areas, one set per pipelined slot: the kernels of two consecutive chunks of a host-buffer call run
    // on two streams and overlap on the device, and both index their scratch by blockIdx.x
    if (var.gmem || use_2x) CUC(cudaMalloc(&d->d_scratch, 2 * (size_t)d->grid * N * sizeof(float2)));
    if (NT > 1) CUC(cudaMalloc(&d->d_xsave, 2 * (size_t)d->grid * N * sizeof(float2)));
    for (auto &s : d->slot) {
        CUC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CUC(cudaMalloc(&s.d_in, (size_t)cfg->max_batch * 2 * N));
        CUC(cudaMalloc(&s.d_idx, (size_t)cfg->max_batch * sizeof(int64_t)));
        CUC(cudaMalloc(&s.d_out, (size_t)cfg->max_batch * NT * sizeof(thr_record)));
        CUC(cudaMallocHost(&s.h_out, (size_t)cfg->max_batch * NT * sizeof(thr_record)));
        CUC(cudaMallocHost(&s.h_idx, (size_t)cfg->max_batch * 2 * sizeof(int64_t)));
    }
    d->c64_chunk = cfg->max_batch < 512 ? cfg->max_batch : 512;
    d->host_chunk = d->grid * 4 < 256 ? 256 : d->grid * 4;
    if (d->host_chunk > cfg->max_batch) d->host_chunk = cfg->max_batch;

    // ---- constant kernel parameters
    DetectParams &p = d->base;
    std::memset(&p, 0, sizeof p);
    p.n_templates = NT;
    p.tpl_spec = d->d_tpl;
    p.tpl_energy = d->d_tpl_energy;
    p.tpl_shift = d->d_tpl_shift;
    p.tpl_nat = d->d_tpl_nat;
    p.scratch = d->d_scratch;
    p.xsave = d->d_xsave;
    p.win_start = ws % N;
    p.win_len = (we - ws + 1) > N ? N : (we - ws + 1);
    // pruned FFT#1: every window bin and its +-3 fit neighbours inside [0,128), no stddev term
    p.zoom = zoom_cfg ? 1 : 0;
    p.zoom_base = zoom_cfg ? zoom_base : 0;
    p.zoom_w0 = ((ws % N) - p.zoom_base + N) % N;
    p.c_const = (float)cfg->carrier_thresh[0];
    p.c_snr = (float)cfg->carrier_thresh[1];
    p.c_std = (float)cfg->carrier_thresh[2];
    p.k_const = (float)cfg->corr_thresh[0];
    p.k_snr = (float)cfg->corr_thresh[1];
    p.k_std = (float)cfg->corr_thresh[2];
    {   // soa_estimator.py:20-39 calculate_window
        const int corr_len = N - L + 1, padding = H - L + 1, left = padding / 2, right = padding - left;
        p.corr_len = corr_len;
        p.corr_start = left;
        p.corr_stop = corr_len - right;
    }
    p.new_len = N - H;
    {
        p.fit_W = (double)cfg->carrier_len;
        p.fit_piW = 3.141592653589793 * (double)cfg->carrier_len;
        p.fit_N = (double)N;
    }
    *out = d;
    return THR_OK;
#undef CUC
}

int thr_get_info(const thr_detector *d, thr_info *info) {
    if (!d || !info) return THR_ERR_INVALID;
    std::memset(info, 0, sizeof *info);
    info->abi_version = THR_ABI_VERSION;
    info->device = d->device;
    info->sm_count = d->sm_count;
    info->grid = d->grid;
    info->threads = d->var.launch_threads;
    info->smem_bytes = (int32_t)d->var.smem;
    info->ctas_per_sm = d->ctas_per_sm;
    info->buffer_in_smem = d->var.gmem ? 0 : 1;
    info->launches = d->launches;
    std::snprintf(info->device_name, sizeof info->device_name, "%s", d->device_name);
    std::snprintf(info->kernel, sizeof info->kernel, "%s", d->var.name);
    return THR_OK;
}

int thr_set_stream(thr_detector *d, void *cuda_stream) {
    if (!d) return THR_ERR_INVALID;
    d->stream = cuda_stream ? (cudaStream_t)cuda_stream : d->own_stream;
    return THR_OK;
}

int thr_synchronize(thr_detector *d) {
    if (!d) return THR_ERR_INVALID;
    CU(d, cudaSetDevice(d->device));
    CU(d, cudaStreamSynchronize(d->stream));
    for (auto &s : d->slot) CU(d, cudaStreamSynchronize(s.stream));
    return THR_OK;
}

int thr_timer_start(thr_detector *d) {
    if (!d) return THR_ERR_INVALID;
    CU(d, cudaSetDevice(d->device));
    CU(d, cudaEventRecord(d->ev0, d->stream));
    return THR_OK;
}

int thr_timer_stop(thr_detector *d, float *ms) {
    if (!d || !ms) return THR_ERR_INVALID;
    CU(d, cudaSetDevice(d->device));
    CU(d, cudaEventRecord(d->ev1, d->stream));
    CU(d, cudaEventSynchronize(d->ev1));
    CU(d, cudaEventElapsedTime(ms, d->ev0, d->ev1));
    return THR_OK;
}

int thr_detect_batch_device(thr_detector *d, const uint8_t *d_raw, const int64_t *d_idx, int32_t n_blocks,
                            thr_record *d_out) {
    if (!d || !d_raw || !d_out || n_blocks < 0) return d ? fail(d, THR_ERR_INVALID, "null/negative argument") : THR_ERR_INVALID;
    if (n_blocks > d->cfg.max_batch) return fail(d, THR_ERR_INVALID, "n_blocks %d exceeds max_batch %d", n_blocks, d->cfg.max_batch);
    if (((uintptr_t)d_raw & 15) != 0) return fail(d, THR_ERR_INVALID, "d_raw must be 16-byte aligned (TMA bulk copy)");
    CU(d, cudaSetDevice(d->device));
    return launch(d, d->stream, d_raw, nullptr, d_idx, n_blocks, d_out, nullptr, nullptr, nullptr, true);
}

int thr_detect_batch_device_c64(thr_detector *d, const float *d_iq, const int64_t *d_idx, int32_t n_blocks,
                                thr_record *d_out) {
    if (!d || !d_iq || !d_out || n_blocks < 0) return d ? fail(d, THR_ERR_INVALID, "null/negative argument") : THR_ERR_INVALID;
    if (n_blocks > d->cfg.max_batch) return fail(d, THR_ERR_INVALID, "n_blocks %d exceeds max_batch %d", n_blocks, d->cfg.max_batch);
    if (((uintptr_t)d_iq & 7) != 0) return fail(d, THR_ERR_INVALID, "d_iq must be 8-byte aligned");
    CU(d, cudaSetDevice(d->device));
    return launch(d, d->stream, nullptr, d_iq, d_idx, n_blocks, d_out, nullptr, nullptr, nullptr, true);
}

// memcpy on several threads: one core moves ~10 GB/s, PCIe Gen5 takes 50.  THRIFTY_B200_COPY_THREADS (1..16) overrides
// the default of 8 threads (4 on hosts with fewer than 16 hardware threads).
static int copy_threads() {
    static const int n = [] {
        int v = std::thread::hardware_concurrency() >= 16 ? 8 : 4;
        if (const char *e = std::getenv("THRIFTY_B200_COPY_THREADS")) v = std::atoi(e);
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    return n;
}
// Copy into page-locked staging with non-temporal stores: the destination is read next by the DMA engine, never by this
// core, so a regular memcpy's read-for-ownership of every destination line is wasted DRAM traffic (3 streams instead of 2).
// THRIFTY_B200_COPY_NT=0 selects plain memcpy.
static void stream_copy(char *dst, const char *src, size_t len) {
#if defined(__SSE2__)
    static const bool nt = [] {
        const char *e = std::getenv("THRIFTY_B200_COPY_NT");
        return !(e && e[0] == '0');
    }();
    if (nt && len >= 4096) {
        const size_t head = (64 - ((uintptr_t)dst & 63)) & 63;
        std::memcpy(dst, src, head);
        dst += head, src += head, len -= head;
        const size_t lines = len / 64;
        for (size_t i = 0; i < lines; ++i) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(src) + 0), b = _mm_loadu_si128((const __m128i *)(src) + 1);
            const __m128i c = _mm_loadu_si128((const __m128i *)(src) + 2), e = _mm_loadu_si128((const __m128i *)(src) + 3);
            _mm_stream_si128((__m128i *)(dst) + 0, a);
            _mm_stream_si128((__m128i *)(dst) + 1, b);
            _mm_stream_si128((__m128i *)(dst) + 2, c);
            _mm_stream_si128((__m128i *)(dst) + 3, e);
            src += 64, dst += 64;
        }
        _mm_sfence();
        len -= lines * 64;
    }
#endif
    std::memcpy(dst, src, len);
}
static void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    constexpr size_t MIN_PART = (size_t)2 << 20;
    int parts = (int)(bytes / MIN_PART);
    if (parts > copy_threads()) parts = copy_threads();
    if (parts <= 1) {
        stream_copy((char *)dst, (const char *)src, bytes);
        return;
    }
    const size_t part = (bytes / parts) & ~(size_t)63;            // the last part also takes the remainder
    std::thread th[15];
    for (int i = 1; i < parts; ++i) {
        const size_t off = (size_t)i * part, len = (i == parts - 1) ? bytes - off : part;
        th[i - 1] = std::thread([=] { stream_copy((char *)dst + off, (const char *)src + off, len); });
    }
    stream_copy((char *)dst, (const char *)src, part);
    for (int i = 1; i < parts; ++i) th[i - 1].join();
}

static bool is_pageable(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();                        // unregistered host memory reports an error on old drivers
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// Pageable input (a NumPy array, a Python bytes object): an "asynchronous" copy from it is staged by the driver at
// ~10 GB/s and blocks the calling thread.  Copying the chunk into page-locked staging ourselves (several threads) runs at
// 25+ GB/s and leaves the DMA truly asynchronous, so it overlaps the kernel of the other slot.
static int stage_pageable(thr_detector *d, Slot &s, const void **src, size_t bytes, size_t cap_hint) {
    NvtxRange r("thr: pageable input -> page-locked staging");
    if (s.h_in_cap < bytes) {
        if (s.h_in) cudaFreeHost(s.h_in);
        s.h_in = nullptr;
        s.h_in_cap = cap_hint > bytes ? cap_hint : bytes;
        CU(d, cudaMallocHost(&s.h_in, s.h_in_cap));
    }
    parallel_memcpy(s.h_in, *src, bytes);
    *src = s.h_in;
    return THR_OK;
}

// Slot hand-over of the host-buffer entry points: wait for the chunk that used this slot two chunks ago and move
// its records from the page-locked staging buffer to the caller's buffer.
static int slot_flush(thr_detector *d, Slot &s) {
    CU(d, cudaStreamSynchronize(s.stream));
    if (s.pending_n) {
        std::memcpy(s.pending_dst, s.h_out, s.pending_n * sizeof(thr_record));
        s.pending_n = 0;
    }
    return THR_OK;
}
// Entry of every host-buffer call: a previous call that failed half-way may have left records queued for a buffer the
// caller no longer owns -- drop them (after the streams have drained) instead of copying into it.
static int slots_reset(thr_detector *d) {
    for (auto &s : d->slot) {
        CU(d, cudaStreamSynchronize(s.stream));
        s.pending_n = 0;
        s.pending_dst = nullptr;
    }
    return THR_OK;
}
static int slot_queue_records(thr_detector *d, Slot &s, thr_record *dst, size_t n) {
    CU(d, cudaMemcpyAsync(s.h_out, s.d_out, n * sizeof(thr_record), cudaMemcpyDeviceToHost, s.stream));
    s.pending_dst = dst;
    s.pending_n = n;
    return THR_OK;
}

// cursor != NULL: the chunks of this call are drawn from a counter shared with other handles (thr_group_detect_batch: the
// GPUs of a group take chunks of one batch as fast as their PCIe paths deliver them); records still land at their
// block's position in `out`.
static int detect_host(thr_detector *d, const uint8_t *raw, const float *iq, const int64_t *block_idx,
                       int64_t n_blocks, thr_record *out, std::atomic<int64_t> *cursor = nullptr) {
    if (!d || (!raw && !iq) || !out || n_blocks < 0) return d ? fail(d, THR_ERR_INVALID, "null/negative argument") : THR_ERR_INVALID;
    CU(d, cudaSetDevice(d->device));
    const int N = d->cfg.block_len, NT = d->cfg.n_templates;
    // chunks small enough that the H2D copy of chunk c+1 overlaps the kernel of chunk c even within
    // one max_batch-sized call, large enough to give every persistent CTA a few blocks
    int64_t chunk = raw ? d->host_chunk : d->c64_chunk;
    if (iq) {
        for (auto &s : d->slot)
            if (!s.d_iq) CU(d, cudaMalloc(&s.d_iq, (size_t)d->c64_chunk * N * 8));
    }
    const bool pageable = is_pageable(raw ? (const void *)raw : (const void *)iq);
    if (int rc0 = slots_reset(d)) return rc0;
    // A call ends with the link idle while the last chunk's kernel and record copy drain; so the tail of a call is cut
    // into ever smaller chunks (half of what is left, down to two blocks per CTA): the last kernel is short.
    auto chunk_at = [&](int64_t b0) -> int64_t {
        if (cursor) return chunk;
        const int64_t left = n_blocks - b0;
        if (left > chunk) return chunk;
        const int64_t small = 2 * (int64_t)d->grid;
        if (left <= small) return left;
        const int64_t half = ((left / 2 + d->grid - 1) / d->grid) * d->grid;
        return half < small ? small : half;
    };
    int c = 0;
    int64_t this_chunk = chunk_at(0);
    for (int64_t b0 = cursor ? cursor->fetch_add(chunk) : 0; b0 < n_blocks;
         b0 = cursor ? cursor->fetch_add(chunk) : b0 + this_chunk, this_chunk = chunk_at(b0), ++c) {
        Slot &s = d->slot[c & 1];
        const int nb = (int)((n_blocks - b0) < this_chunk ? (n_blocks - b0) : this_chunk);
        int rc;
        {
            NvtxRange r("thr: records of chunk c-2 -> caller");
            rc = slot_flush(d, s);                 // chunk c-2 done: its records go to the caller, staging is free
        }
        if (rc != THR_OK) return rc;
        NvtxRange r_chunk("thr: chunk (stage, H2D, launch, D2H queued)");
        const void *src = raw ? (const void *)(raw + (size_t)b0 * 2 * N) : (const void *)(iq + (size_t)b0 * 2 * N);
        const size_t bytes = raw ? (size_t)nb * 2 * N : (size_t)nb * N * 8;
        if (pageable) {
            rc = stage_pageable(d, s, &src, bytes, (size_t)chunk * (raw ? 2 * N : 8 * N));
            if (rc != THR_OK) return rc;
        }
        CU(d, cudaMemcpyAsync(raw ? (void *)s.d_in : (void *)s.d_iq, src, bytes, cudaMemcpyHostToDevice, s.stream));
        for (int i = 0; i < nb; ++i) s.h_idx[i] = block_idx ? block_idx[b0 + i] : b0 + i;
        CU(d, cudaMemcpyAsync(s.d_idx, s.h_idx, (size_t)nb * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        rc = launch(d, s.stream, raw ? s.d_in : nullptr, raw ? nullptr : s.d_iq, s.d_idx, nb, s.d_out, nullptr,
                    nullptr, nullptr);
        if (rc != THR_OK) return rc;
        rc = slot_queue_records(d, s, out + (size_t)b0 * NT, (size_t)nb * NT);
        if (rc != THR_OK) return rc;
    }
    for (auto &s : d->slot) {
        int rc = slot_flush(d, s);
        if (rc != THR_OK) return rc;
    }
    CU(d, cudaGetLastError());
    return THR_OK;
}

int thr_detect_batch(thr_detector *d, const uint8_t *raw, const int64_t *block_idx, int64_t n_blocks,
                     thr_record *out) {
    NvtxRange r("thr_detect_batch");
    if (!raw) return d ? fail(d, THR_ERR_INVALID, "raw is NULL") : THR_ERR_INVALID;
    return detect_host(d, raw, nullptr, block_idx, n_blocks, out);
}

int thr_detect_batch_c64(thr_detector *d, const float *iq, const int64_t *block_idx, int64_t n_blocks,
                         thr_record *out) {
    if (!iq) return d ? fail(d, THR_ERR_INVALID, "iq is NULL") : THR_ERR_INVALID;
    return detect_host(d, nullptr, iq, block_idx, n_blocks, out);
}

int thr_detect_block_data(thr_detector *d, const uint8_t *raw, const float *iq, int64_t block_idx, thr_record *out,
                          float *shifted_fft, float *corr, float *fft_mag) {
    if (!d || (!raw && !iq) || !out) return d ? fail(d, THR_ERR_INVALID, "null argument") : THR_ERR_INVALID;
    CU(d, cudaSetDevice(d->device));
    const int N = d->cfg.block_len, NT = d->cfg.n_templates, L = d->cfg.template_len;
    const int corr_len = N - L + 1;
    Slot &s = d->slot[0];
    float2 *d_sfft = nullptr, *d_corr = nullptr;
    float *d_mag = nullptr;
    int rc = THR_OK;
    cudaError_t e = cudaSuccess;
    auto done = [&](int code) {
        cudaFree(d_sfft);
        cudaFree(d_corr);
        cudaFree(d_mag);
        return code;
    };
#define CUD(call)                                                                                    \
    do {                                                                                             \
        e = (call);                                                                                  \
        if (e != cudaSuccess) return done(fail(d, THR_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e))); \
    } while (0)
    CUD(cudaMalloc(&d_sfft, (size_t)N * 8));
    CUD(cudaMalloc(&d_corr, (size_t)corr_len * 8));
    CUD(cudaMalloc(&d_mag, (size_t)N * 4));
    CUD(cudaMemsetAsync(d_sfft, 0, (size_t)N * 8, s.stream));
    CUD(cudaMemsetAsync(d_corr, 0, (size_t)corr_len * 8, s.stream));
    if (iq && !s.d_iq) CUD(cudaMalloc(&s.d_iq, (size_t)d->c64_chunk * N * 8));
    if (raw) CUD(cudaMemcpyAsync(s.d_in, raw, (size_t)2 * N, cudaMemcpyHostToDevice, s.stream));
    else CUD(cudaMemcpyAsync(s.d_iq, iq, (size_t)N * 8, cudaMemcpyHostToDevice, s.stream));
    CUD(cudaMemcpyAsync(s.d_idx, &block_idx, sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
    rc = launch(d, s.stream, raw ? s.d_in : nullptr, raw ? nullptr : s.d_iq, s.d_idx, 1, s.d_out, d_sfft, d_corr, d_mag);
    if (rc != THR_OK) return done(rc);
    CUD(cudaMemcpyAsync(out, s.d_out, (size_t)NT * sizeof(thr_record), cudaMemcpyDeviceToHost, s.stream));
    if (shifted_fft) CUD(cudaMemcpyAsync(shifted_fft, d_sfft, (size_t)N * 8, cudaMemcpyDeviceToHost, s.stream));
    if (corr) CUD(cudaMemcpyAsync(corr, d_corr, (size_t)corr_len * 8, cudaMemcpyDeviceToHost, s.stream));
    if (fft_mag) CUD(cudaMemcpyAsync(fft_mag, d_mag, (size_t)N * 4, cudaMemcpyDeviceToHost, s.stream));
    CUD(cudaStreamSynchronize(s.stream));
    CUD(cudaGetLastError());
    return done(THR_OK);
#undef CUD
}

// ---- stage boundaries of the reference's plug-in seam ---------------------------------------------------------
// thrifty/carrier_sync.py:52-76 Synchronizer.sync: block -> (shifted_fft, CarrierSyncInfo);
// thrifty/soa_estimator.py:78-92 SoaEstimator.soa_estimate: shifted_fft -> (detected, CorrDetectionInfo, corr).
static int stage_setup(thr_detector *d, int st) {
    const int N = d->cfg.block_len;
    if (!d->grid_stage[st]) {
        Variant v;
        if (!thr::pick_variant_stage(N, st, &v))
            return fail(d, THR_ERR_INVALID, "no stage-boundary kernel for block_len %d in this build", N);
        CU(d, cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
        int occ = 0;
        CU(d, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.fn, v.launch_threads, v.smem));
        if (occ < 1) return fail(d, THR_ERR_CUDA, "kernel %s does not fit on an SM (smem %zu)", v.name, v.smem);
        int grid = d->sm_count * occ;
        if (v.gmem) {                                            // global-scratch kernel (block_len 32768)
            if (!d->d_scratch) CU(d, cudaMalloc(&d->d_scratch, 2 * (size_t)d->grid * N * sizeof(float2)));
            d->base.scratch = d->d_scratch;
            if (grid > d->grid) grid = d->grid;                  // the scratch is sized for the main kernel's grid
        }
        if (const char *cap = std::getenv("THRIFTY_B200_MAX_GRID")) {
            const int g = std::atoi(cap);
            if (g >= 1 && g < grid) grid = g;
        }
        d->var_stage[st] = v;
        d->grid_stage[st] = grid;
    }
    if (!d->stage_chunk) {
        int64_t c = ((int64_t)64 << 20) / ((int64_t)N * 8);      // 64 MiB of spectra per chunk at most
        if (c > d->cfg.max_batch) c = d->cfg.max_batch;
        if (c < 1) c = 1;
        d->stage_chunk = (int)c;
    }
    if (!d->d_stage_fft) CU(d, cudaMalloc(&d->d_stage_fft, (size_t)d->stage_chunk * N * sizeof(float2)));
    return THR_OK;
}

static int launch_stage(thr_detector *d, int st, cudaStream_t stream, const uint8_t *d_raw, const float *d_iq,
                        const int64_t *d_idx, int nb, thr_record *d_out, bool want_corr) {
    const int N = d->cfg.block_len, corr_len = N - d->cfg.template_len + 1;
    DetectParams p = d->base;
    p.raw = d_raw;
    p.raw_stride = 2 * (int64_t)N;
    p.iq = reinterpret_cast<const float2 *>(d_iq);
    p.block_idx = d_idx;
    p.out = d_out;
    p.n_blocks = nb;
    p.n_templates = 1;                                           // the stage kernels are one-template kernels
    if (d->has_generic) p.tpl_spec = d->d_tpl_generic;           // 2 x 16384 detectors: the generic kernel's template order
    if (st == 1) {
        p.dbg_shifted_fft = d->d_stage_fft;
        p.dbg_sfft_stride = N;
    } else {
        p.in_sfft = d->d_stage_fft;
        p.dbg_corr = want_corr ? d->d_stage_corr : nullptr;
        p.dbg_corr_stride = corr_len;
    }
    const Variant &v = d->var_stage[st];
    void *args[] = {&p};
    cudaLaunchConfig_t lc;
    std::memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(nb < d->grid_stage[st] ? nb : d->grid_stage[st]);
    lc.blockDim = dim3(v.launch_threads);
    lc.dynamicSmemBytes = v.smem;
    lc.stream = stream;
    CU(d, cudaLaunchKernelExC(&lc, v.fn, args));
    d->launches++;
    return THR_OK;
}

int thr_sync_batch(thr_detector *d, const uint8_t *raw, const float *iq, const int64_t *block_idx, int64_t n_blocks,
                   thr_record *out, float *shifted_fft) {
    if (!d || (!raw && !iq) || !out || !shifted_fft || n_blocks < 0)
        return d ? fail(d, THR_ERR_INVALID, "null/negative argument") : THR_ERR_INVALID;
    if (d->cfg.flags & THR_CFG_FASTDET_SEMANTICS) return fail(d, THR_ERR_INVALID, "thr_sync_batch follows the Python path's semantics");
    CU(d, cudaSetDevice(d->device));
    if (int rc = stage_setup(d, 1)) return rc;
    if (int rc = slots_reset(d)) return rc;
    const int N = d->cfg.block_len;
    Slot &s = d->slot[0];
    if (iq && !s.d_iq) CU(d, cudaMalloc(&s.d_iq, (size_t)d->c64_chunk * N * 8));
    const int64_t chunk = iq ? (d->stage_chunk < d->c64_chunk ? d->stage_chunk : d->c64_chunk) : d->stage_chunk;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += chunk) {
        const int nb = (int)((n_blocks - b0) < chunk ? (n_blocks - b0) : chunk);
        if (raw) CU(d, cudaMemcpyAsync(s.d_in, raw + (size_t)b0 * 2 * N, (size_t)nb * 2 * N, cudaMemcpyHostToDevice, s.stream));
        else CU(d, cudaMemcpyAsync(s.d_iq, iq + (size_t)b0 * 2 * N, (size_t)nb * N * 8, cudaMemcpyHostToDevice, s.stream));
        for (int i = 0; i < nb; ++i) s.h_idx[i] = block_idx ? block_idx[b0 + i] : b0 + i;
        CU(d, cudaMemcpyAsync(s.d_idx, s.h_idx, (size_t)nb * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        CU(d, cudaMemsetAsync(d->d_stage_fft, 0, (size_t)nb * N * sizeof(float2), s.stream));   // rows without a carrier: zeros
        if (int rc = launch_stage(d, 1, s.stream, raw ? s.d_in : nullptr, raw ? nullptr : s.d_iq, s.d_idx, nb, s.d_out, false)) return rc;
        CU(d, cudaMemcpyAsync(out + (size_t)b0, s.d_out, (size_t)nb * sizeof(thr_record), cudaMemcpyDeviceToHost, s.stream));
        CU(d, cudaMemcpyAsync(shifted_fft + (size_t)b0 * 2 * N, d->d_stage_fft, (size_t)nb * N * sizeof(float2),
                              cudaMemcpyDeviceToHost, s.stream));
        CU(d, cudaStreamSynchronize(s.stream));
    }
    CU(d, cudaGetLastError());
    return THR_OK;
}

int thr_soa_batch(thr_detector *d, const float *fft, const int64_t *block_idx, int64_t n_blocks, thr_record *out,
                  float *corr) {
    if (!d || !fft || !out || n_blocks < 0) return d ? fail(d, THR_ERR_INVALID, "null/negative argument") : THR_ERR_INVALID;
    if (d->cfg.flags & THR_CFG_FASTDET_SEMANTICS) return fail(d, THR_ERR_INVALID, "thr_soa_batch follows the Python path's semantics");
    if (d->cfg.n_templates != 1) return fail(d, THR_ERR_INVALID, "thr_soa_batch takes a one-template detector");
    CU(d, cudaSetDevice(d->device));
    if (int rc = stage_setup(d, 2)) return rc;
    if (int rc = slots_reset(d)) return rc;
    const int N = d->cfg.block_len, corr_len = N - d->cfg.template_len + 1;
    if (corr && !d->d_stage_corr) CU(d, cudaMalloc(&d->d_stage_corr, (size_t)d->stage_chunk * corr_len * sizeof(float2)));
    Slot &s = d->slot[0];
    const int64_t chunk = d->stage_chunk;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += chunk) {
        const int nb = (int)((n_blocks - b0) < chunk ? (n_blocks - b0) : chunk);
        CU(d, cudaMemcpyAsync(d->d_stage_fft, fft + (size_t)b0 * 2 * N, (size_t)nb * N * sizeof(float2), cudaMemcpyHostToDevice, s.stream));
        for (int i = 0; i < nb; ++i) s.h_idx[i] = block_idx ? block_idx[b0 + i] : b0 + i;
        CU(d, cudaMemcpyAsync(s.d_idx, s.h_idx, (size_t)nb * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        if (int rc = launch_stage(d, 2, s.stream, nullptr, nullptr, s.d_idx, nb, s.d_out, corr != nullptr)) return rc;
        CU(d, cudaMemcpyAsync(out + (size_t)b0, s.d_out, (size_t)nb * sizeof(thr_record), cudaMemcpyDeviceToHost, s.stream));
        if (corr)
            CU(d, cudaMemcpyAsync(corr + (size_t)b0 * 2 * corr_len, d->d_stage_corr, (size_t)nb * corr_len * sizeof(float2),
                                  cudaMemcpyDeviceToHost, s.stream));
        CU(d, cudaStreamSynchronize(s.stream));
    }
    CU(d, cudaGetLastError());
    return THR_OK;
}

// ---- .card text ingest ------------------------------------------------------------------
// thrifty/block_data.py:101-131 card_reader / fastcard/card_reader.c:22-78
static bool starts_with(const char *p, const char *e, const char *lit) {
    const size_t n = std::strlen(lit);
    return (size_t)(e - p) >= n && std::memcmp(p, lit, n) == 0;
}

// `lines` (optional): in = lines seen by earlier calls on the same text, out = lines seen including this call, so that
// a caller scanning piecewise reports the same line numbers as one scan over the whole text.
static int card_scan_lines(const char *text, size_t len, int32_t block_len, int32_t final_chunk, int64_t max_blocks,
                           double *timestamps, int64_t *block_idx, int64_t *payload_off, int64_t *n_found,
                           int64_t *consumed, int64_t *bad_line, int64_t *lines) {
    if (!text || !timestamps || !block_idx || !payload_off || !n_found || !consumed || block_len < 1)
        return THR_ERR_INVALID;
    const int64_t want = ((2 * (int64_t)block_len + 2) / 3) * 4;       // base64 characters per payload
    const char *p = text, *end = text + len;
    int64_t n = 0, line_no = lines ? *lines : 0;
    if (bad_line) *bad_line = -1;
    while (p < end && n < max_blocks) {
        // Fast path: a data line's payload has a known length, so after the two numbers the end of the
        // line is found by a jump instead of a scan over 43.7 KB of base64 text.
        // (the number parsers stop at the first character that is not part of a number: a space inside the next 40
        // bytes guarantees that they stay inside [p, end) even when the text is not NUL-terminated)
        const size_t look = (size_t)(end - p) < 40 ? (size_t)(end - p) : 40;
        if (*p != '#' && *p != '\n' && *p != '\r' && *p != 'U' && *p != 'l' && std::memchr(p, ' ', look)) {
            char *q = nullptr;
            const double ts = std::strtod(p, &q);
            if (q != p && q < end && *q == ' ') {
                const char *r = q + 1;
                char *q2 = nullptr;
                const long long idx = std::strtoll(r, &q2, 10);
                if (q2 != r && q2 < end && *q2 == ' ') {
                    const char *pay = q2 + 1, *pe = pay + want;
                    const char *next = nullptr;
                    if (pe < end && *pe == '\n') next = pe + 1;
                    else if (pe + 1 < end && pe[0] == '\r' && pe[1] == '\n') next = pe + 2;
                    else if (pe == end && final_chunk) next = pe;
                    if (next) {
                        ++line_no;
                        timestamps[n] = ts;
                        block_idx[n] = idx;
                        payload_off[n] = pay - text;
                        ++n;
                        p = next;
                        continue;
                    }
                }
            }
        }
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        if (!nl && !final_chunk) break;                                // incomplete last line: wait for more
        const char *e = nl ? nl : end;
        const char *next = nl ? nl + 1 : end;
        ++line_no;
        const char *le = e;
        if (le > p && le[-1] == '\r') --le;
        if (le == p || *p == '#' || starts_with(p, le, "Using Volk machine:") || starts_with(p, le, "linux;")) {
            p = next;                                                  // comment / blank / stdout noise
            continue;
        }
        char *q = nullptr;
        if (!std::memchr(p, ' ', (size_t)(le - p))) {                  // no separator at all: malformed, and strtod must
            if (bad_line) *bad_line = line_no;                         // not run off an unterminated buffer
            *n_found = n;
            *consumed = p - text;
            return THR_ERR_INVALID;
        }
        const double ts = std::strtod(p, &q);
        bool ok = q != p && q < le && *q == ' ';
        long long idx = 0;
        if (ok) {
            const char *r = q + 1;
            char *q2 = nullptr;
            idx = std::strtoll(r, &q2, 10);
            ok = q2 != r && q2 < le && *q2 == ' ';
            q = q2;
        }
        if (ok) ok = (le - (q + 1)) == want;                           // card_reader.c:58-66 length check
        if (!ok) {
            if (bad_line) *bad_line = line_no;
            *n_found = n;
            *consumed = p - text;
            return THR_ERR_INVALID;
        }
        timestamps[n] = ts;
        block_idx[n] = idx;
        payload_off[n] = (q + 1) - text;
        ++n;
        p = next;
    }
    *n_found = n;
    *consumed = p - text;
    if (lines) *lines = line_no;
    return THR_OK;
}

int thr_card_scan(const char *text, size_t len, int32_t block_len, int32_t final_chunk, int64_t max_blocks,
                  double *timestamps, int64_t *block_idx, int64_t *payload_off, int64_t *n_found,
                  int64_t *consumed, int64_t *bad_line) {
    return card_scan_lines(text, len, block_len, final_chunk, max_blocks, timestamps, block_idx, payload_off, n_found,
                           consumed, bad_line, nullptr);
}

int thr_detect_card(thr_detector *d, const char *text, size_t len, int32_t final_chunk, int64_t max_blocks,
                    double *timestamps, int64_t *block_idx, thr_record *out, int64_t *n_blocks,
                    int64_t *consumed) {
    if (!d || !text || !timestamps || !block_idx || !out || !n_blocks || !consumed)
        return d ? fail(d, THR_ERR_INVALID, "null argument") : THR_ERR_INVALID;
    NvtxRange r_call("thr_detect_card");
    CU(d, cudaSetDevice(d->device));
    const int N = d->cfg.block_len, NT = d->cfg.n_templates;
    const int64_t want = ((2 * (int64_t)N + 2) / 3) * 4;
    if (max_blocks < 0) return fail(d, THR_ERR_INVALID, "max_blocks is negative");
    *n_blocks = 0;
    *consumed = 0;
    if (int rc0 = slots_reset(d)) return rc0;
    if (!d->d_bad) {
        CU(d, cudaMalloc(&d->d_bad, 8 * sizeof(unsigned int)));
    }
    CU(d, cudaMemsetAsync(d->d_bad, 0, 8 * sizeof(unsigned int), d->slot[0].stream));
    CU(d, cudaStreamSynchronize(d->slot[0].stream));
    const int64_t chunk = d->host_chunk;
    const bool pageable = is_pageable(text);
    std::vector<int64_t> off((size_t)chunk);
    // The lines of chunk c+1 are scanned while chunk c crosses PCIe: each line costs two cache (and TLB) misses in a text of
    // hundreds of megabytes, ~1 ms per 4096 lines if done up front.
    int64_t pos = 0, b0 = 0, lines = 0;
    int rc = THR_OK;
    for (int c = 0; b0 < max_blocks; ++c) {
        const int64_t ask = (max_blocks - b0) < chunk ? (max_blocks - b0) : chunk;
        int64_t got = 0, used = 0, bad_line = -1;
        {
            NvtxRange r("thr: scan .card line headers");
            rc = card_scan_lines(text + pos, len - (size_t)pos, N, final_chunk, ask, timestamps + b0, block_idx + b0, off.data(),
                                 &got, &used, &bad_line, &lines);
        }
        if (rc != THR_OK)
            return fail(d, rc, ".card data line %lld is malformed (expected '<time> <index> <%lld base64 chars>')",
                        (long long)bad_line, (long long)want);
        const char *ctext = text + pos;            // payload offsets of this chunk are relative to it
        pos += used;
        *consumed = pos;
        if (got == 0) break;
        Slot &s = d->slot[c & 1];
        const int nb = (int)got;
        rc = slot_flush(d, s);
        if (rc != THR_OK) return rc;
        // byte range of the text that covers the payloads of this chunk (aligned down to 4 in the caller's buffer)
        const int64_t lead = (int64_t)((uintptr_t)(ctext + off[0]) & 3);
        const int64_t t0 = off[0] - lead, t1 = off[nb - 1] + want;
        const size_t bytes = (size_t)(t1 - t0);
        if (s.text_cap < bytes + 16) {
            cudaFree(s.d_text);
            s.d_text = nullptr;
            s.text_cap = bytes + bytes / 4 + 4096;
            CU(d, cudaMalloc(&s.d_text, s.text_cap));
        }
        if (!s.d_off) CU(d, cudaMalloc(&s.d_off, (size_t)d->cfg.max_batch * sizeof(int64_t)));
        int64_t *h_rel = s.h_idx + d->cfg.max_batch;
        for (int i = 0; i < nb; ++i) {
            s.h_idx[i] = block_idx[b0 + i];
            h_rel[i] = off[i] - t0;
        }
        const void *tsrc = ctext + t0;
        if (pageable) {
            rc = stage_pageable(d, s, &tsrc, bytes, bytes + bytes / 4);
            if (rc != THR_OK) return rc;
        }
        CU(d, cudaMemcpyAsync(s.d_text, tsrc, bytes, cudaMemcpyHostToDevice, s.stream));
        CU(d, cudaMemcpyAsync(s.d_off, h_rel, (size_t)nb * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        CU(d, cudaMemcpyAsync(s.d_idx, s.h_idx, (size_t)nb * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        const dim3 grid((unsigned)((want + thr::B64_SEG_CHARS - 1) / thr::B64_SEG_CHARS), (unsigned)nb);
        thr::b64_decode_kernel<<<grid, thr::B64_THREADS, 0, s.stream>>>(s.d_text, s.d_off, (int)want, 2 * N, s.d_in,
                                                                       d->d_bad, (int)b0);
        CU(d, cudaGetLastError());
        d->launches++;
        rc = launch(d, s.stream, s.d_in, nullptr, s.d_idx, nb, s.d_out, nullptr, nullptr, nullptr);
        if (rc != THR_OK) return rc;
        rc = slot_queue_records(d, s, out + (size_t)b0 * NT, (size_t)nb * NT);
        if (rc != THR_OK) return rc;
        b0 += got;
        *n_blocks = b0;
        if (got < ask) break;                      // end of the text (or an unterminated last line) reached
    }
    for (auto &s : d->slot) {
        rc = slot_flush(d, s);
        if (rc != THR_OK) return rc;
    }
    unsigned int n_bad[8] = {0};
    CU(d, cudaMemcpy(n_bad, d->d_bad, sizeof n_bad, cudaMemcpyDeviceToHost));
    if (n_bad[0]) {
        char shown[17];
        for (int i = 0; i < 16; ++i) {
            const unsigned char ch = (unsigned char)(n_bad[3 + i / 4] >> (8 * (i % 4)));
            shown[i] = (ch >= 32 && ch < 127) ? (char)ch : '?';
        }
        shown[16] = 0;
        return fail(d, THR_ERR_INVALID, ".card payload contains %u group(s) with non-base64 characters (first: data line %u of "
                    "this call, payload character %u: \"%s\")", n_bad[0], n_bad[1] + 1, n_bad[2], shown);
    }
    return THR_OK;
}


// ---- raw sample streams (thrifty/block_data.py:70-98 block_reader, fastcard/raw_reader.c:15-46) ------------
// Block b of a contiguous uint8 I/Q stream covers samples [b*(N-H) - H, b*(N-H) + N - H).  The kernel reads the
// overlapping windows straight from the stream (TMA tile source = stream + 2*(b*(N-H) - H)), so the host never
// re-blocks and only N-H new samples per block cross PCIe.  Block 0 (whose history precedes the stream) is the
// caller's business: pass first_block >= 1, or a stream that already starts with H samples of history.
int thr_detect_stream_device(thr_detector *d, const uint8_t *d_stream, int64_t n_stream_bytes, const int64_t *d_block_idx,
                             int32_t n_blocks, thr_record *d_out) {
    if (!d || !d_stream || !d_out || n_blocks < 0) return d ? fail(d, THR_ERR_INVALID, "null/negative argument") : THR_ERR_INVALID;
    const int64_t N = d->cfg.block_len, H = d->cfg.history_len, stride = 2 * (N - H);
    if (n_blocks > d->cfg.max_batch) return fail(d, THR_ERR_INVALID, "n_blocks %d exceeds max_batch %d", n_blocks, d->cfg.max_batch);
    if ((stride & 15) != 0 || ((uintptr_t)d_stream & 15) != 0)
        return fail(d, THR_ERR_INVALID, "stream mode needs 2*(block_len-history_len) = %lld and the stream pointer to be multiples of 16 bytes (TMA bulk copy)", (long long)stride);
    if (n_blocks > 0 && (int64_t)(n_blocks - 1) * stride + 2 * N > n_stream_bytes)
        return fail(d, THR_ERR_INVALID, "stream of %lld bytes is too short for %d blocks", (long long)n_stream_bytes, n_blocks);
    CU(d, cudaSetDevice(d->device));
    return launch(d, d->stream, d_stream, nullptr, d_block_idx, n_blocks, d_out, nullptr, nullptr, nullptr, true, stride);
}

// Host stream: `stream` holds the H history samples of block `first_block` followed by new samples
// (i.e. stream[0] is sample first_block*(N-H) - H).  Returns the number of whole blocks processed.
int thr_detect_stream(thr_detector *d, const uint8_t *stream, int64_t n_stream_bytes, int64_t first_block,
                      thr_record *out, int64_t *n_blocks_out) {
    if (!d || !stream || !out || !n_blocks_out) return d ? fail(d, THR_ERR_INVALID, "null argument") : THR_ERR_INVALID;
    NvtxRange r_call("thr_detect_stream");
    CU(d, cudaSetDevice(d->device));
    const int64_t N = d->cfg.block_len, H = d->cfg.history_len, stride = 2 * (N - H);
    const int NT = d->cfg.n_templates;
    if ((stride & 15) != 0) return fail(d, THR_ERR_INVALID, "stream mode needs 2*(block_len-history_len) to be a multiple of 16");
    const int64_t nb_total = n_stream_bytes >= 2 * N ? (n_stream_bytes - 2 * N) / stride + 1 : 0;
    *n_blocks_out = nb_total;
    const int64_t chunk = d->host_chunk;
    const bool pageable = is_pageable(stream);
    if (int rc0 = slots_reset(d)) return rc0;
    int c = 0;
    for (int64_t b0 = 0; b0 < nb_total; b0 += chunk, ++c) {
        Slot &s = d->slot[c & 1];
        const int nb = (int)((nb_total - b0) < chunk ? (nb_total - b0) : chunk);
        int rc = slot_flush(d, s);
        if (rc != THR_OK) return rc;
        const size_t bytes = (size_t)((nb - 1) * stride + 2 * N);        // <= nb * 2N: fits the raw staging buffer
        const void *ssrc = stream + b0 * stride;
        if (pageable) {
            rc = stage_pageable(d, s, &ssrc, bytes, (size_t)chunk * 2 * N);
            if (rc != THR_OK) return rc;
        }
        CU(d, cudaMemcpyAsync(s.d_in, ssrc, bytes, cudaMemcpyHostToDevice, s.stream));
        for (int i = 0; i < nb; ++i) s.h_idx[i] = first_block + b0 + i;
        CU(d, cudaMemcpyAsync(s.d_idx, s.h_idx, (size_t)nb * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        rc = launch(d, s.stream, s.d_in, nullptr, s.d_idx, nb, s.d_out, nullptr, nullptr, nullptr, false, stride);
        if (rc != THR_OK) return rc;
        rc = slot_queue_records(d, s, out + (size_t)b0 * NT, (size_t)nb * NT);
        if (rc != THR_OK) return rc;
    }
    for (auto &s : d->slot) {
        int rc = slot_flush(d, s);
        if (rc != THR_OK) return rc;
    }
    CU(d, cudaGetLastError());
    return THR_OK;
}


// ---- several GPUs behind one handle ---------------------------------------------------------------------------
// The detect path has no cross-block state (thrifty/detect.py:40-58) and every .card block carries its own history, so a
// batch shards into contiguous stripes with no data-path collective: one thr_detector + one host thread per GPU, stripe g
// = blocks [g * ceil(B/G), ...), every stripe's records land in its slice of the caller's array, block order preserved
// (the seam is the loop at thrifty/detect.py:217-223).  Each worker thread first moves to the CPUs of its GPU's NUMA node
// (sysfs; best effort) so that its page-locked staging is node-local.
namespace {

int gpu_numa_node(int dev) {
    char busid[32] = {0};
    if (cudaDeviceGetPCIBusId(busid, (int)sizeof busid, dev) != cudaSuccess) return -1;
    for (char *c = busid; *c; ++c) *c = (char)std::tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + busid + "/numa_node";
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}

bool bind_thread_to_node(int node) {
    if (node < 0) return false;
    char path[128];
    std::snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = std::fopen(path, "r");
    if (!f) return false;
    char buf[2048] = {0};
    const bool ok = std::fgets(buf, sizeof buf, f) != nullptr;
    std::fclose(f);
    if (!ok) return false;
    cpu_set_t set, allowed;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return false;
    int n = 0;
    char *save = nullptr;
    for (char *tok = strtok_r(buf, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {
        int lo = 0, hi = 0;
        const int got = std::sscanf(tok, "%d-%d", &lo, &hi);
        if (got == 1) hi = lo;
        if (got < 1) continue;
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) { CPU_SET(c, &set); ++n; }
    }
    return n > 0 && sched_setaffinity(0, sizeof set, &set) == 0;
}

struct GroupWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int(thr_detector *)> job;
    bool has_job = false, done = false, quit = false;
    int rc = THR_OK;
    thr_detector *det = nullptr;
    int device = 0, numa = -1;
    bool bound = false;
    std::string err;
};

}  // namespace

struct thr_group {
    std::vector<std::unique_ptr<GroupWorker>> w;
    thr_config cfg;
    std::string err;
};

namespace {

void group_worker_main(GroupWorker *w, const thr_config *cfg0) {
    w->numa = gpu_numa_node(w->device);
    w->bound = bind_thread_to_node(w->numa);
    thr_config cfg = *cfg0;
    cfg.device = w->device;
    int rc = thr_create(&cfg, &w->det);
    if (rc != THR_OK) w->err = thr_last_error(nullptr);
    {
        std::lock_guard<std::mutex> lk(w->m);
        w->rc = rc;
        w->done = true;
    }
    w->cv.notify_all();
    for (;;) {
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) break;
        auto job = w->job;
        w->has_job = false;
        lk.unlock();
        const int r = job(w->det);
        lk.lock();
        w->rc = r;
        w->done = true;
        lk.unlock();
        w->cv.notify_all();
    }
    if (w->det) thr_destroy(w->det);
    w->det = nullptr;
}

// run job(g, det) on every worker and wait; returns the first non-zero status
int group_run(thr_group *g, const std::function<int(int, thr_detector *)> &job) {
    const int n = (int)g->w.size();
    for (int i = 0; i < n; ++i) {
        GroupWorker *w = g->w[i].get();
        std::lock_guard<std::mutex> lk(w->m);
        w->job = [i, &job](thr_detector *d) { return job(i, d); };
        w->done = false;
        w->has_job = true;
        w->cv.notify_all();
    }
    int rc = THR_OK;
    for (int i = 0; i < n; ++i) {
        GroupWorker *w = g->w[i].get();
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != THR_OK && rc == THR_OK) {
            rc = w->rc;
            g->err = std::string("device ") + std::to_string(w->device) + ": " + thr_last_error(w->det);
        }
    }
    return rc;
}

}  // namespace

void thr_group_destroy(thr_group *g) {
    if (!g) return;
    for (auto &w : g->w) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
        }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
    }
    delete g;
}

int thr_group_create(const thr_config *cfg, const int32_t *devices, int32_t n_devices, thr_group **out) {
    if (!cfg || !devices || !out || n_devices < 1 || n_devices > 64) return fail(nullptr, THR_ERR_INVALID, "bad group arguments");
    *out = nullptr;
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(nullptr, THR_ERR_INVALID, "device %d listed twice", devices[i]);
    thr_group *g = new thr_group();
    g->cfg = *cfg;
    for (int i = 0; i < n_devices; ++i) {
        g->w.emplace_back(new GroupWorker());
        GroupWorker *w = g->w.back().get();
        w->device = devices[i];
        w->th = std::thread(group_worker_main, w, cfg);        // cfg->templates stays valid until every create has returned
    }
    int rc = THR_OK;
    for (auto &w : g->w) {
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != THR_OK && rc == THR_OK) {
            rc = w->rc;
            g_create_error = "device " + std::to_string(w->device) + ": " + w->err;
        }
    }
    g->cfg.templates = nullptr;
    if (rc != THR_OK) {
        const std::string keep = g_create_error;
        thr_group_destroy(g);
        g_create_error = keep;
        return rc;
    }
    *out = g;
    return THR_OK;
}

const char *thr_group_last_error(const thr_group *g) { return g ? g->err.c_str() : g_create_error.c_str(); }
int thr_group_size(const thr_group *g) { return g ? (int)g->w.size() : 0; }
thr_detector *thr_group_member(thr_group *g, int32_t i) {
    return (g && i >= 0 && i < (int)g->w.size()) ? g->w[i]->det : nullptr;
}
int thr_group_numa_node(const thr_group *g, int32_t i, int32_t *bound) {
    if (!g || i < 0 || i >= (int)g->w.size()) return -1;
    if (bound) *bound = g->w[i]->bound ? 1 : 0;
    return g->w[i]->numa;
}

// stripe g of n items over G devices: [lo, hi)
static inline void stripe_of(int64_t n, int G, int g, int64_t *lo, int64_t *hi) {
    const int64_t per = (n + G - 1) / G;
    *lo = per * g < n ? per * g : n;
    *hi = per * (g + 1) < n ? per * (g + 1) : n;
}

int thr_group_detect_batch(thr_group *g, const uint8_t *raw, const int64_t *block_idx, int64_t n_blocks, thr_record *out) {
    if (!g || !raw || !out || n_blocks < 0) return THR_ERR_INVALID;
    // Contiguous chunks handed out from one counter: a GPU behind a slower PCIe path (or busy with something else) simply
    // takes fewer of them, so the batch finishes when the host link is saturated, not when the slowest stripe is done.
    // Every chunk's records are written at its blocks' positions in `out`: input order, whoever computed them.
    std::atomic<int64_t> cursor(0);
    return group_run(g, [&](int, thr_detector *d) -> int {
        return detect_host(d, raw, nullptr, block_idx, n_blocks, out, &cursor);
    });
}

int thr_group_detect_stream(thr_group *g, const uint8_t *stream, int64_t n_stream_bytes, int64_t first_block,
                            thr_record *out, int64_t *n_blocks_out) {
    if (!g || !stream || !out || !n_blocks_out) return THR_ERR_INVALID;
    const int G = (int)g->w.size();
    const int64_t N = g->cfg.block_len, H = g->cfg.history_len, stride = 2 * (N - H), NT = g->cfg.n_templates;
    const int64_t nb_total = n_stream_bytes >= 2 * N ? (n_stream_bytes - 2 * N) / stride + 1 : 0;
    *n_blocks_out = nb_total;
    // each stripe starts with the H samples of history of its first block (the halo): no exchange between GPUs
    return group_run(g, [&](int i, thr_detector *d) -> int {
        int64_t lo, hi;
        stripe_of(nb_total, G, i, &lo, &hi);
        if (hi <= lo) return THR_OK;
        int64_t got = 0;
        const int64_t bytes = (hi - lo - 1) * stride + 2 * N;
        const int rc = thr_detect_stream(d, stream + lo * stride, bytes, first_block + lo, out + (size_t)lo * NT, &got);
        if (rc == THR_OK && got != hi - lo) return fail(d, THR_ERR_INVALID, "stripe of %lld blocks returned %lld", (long long)(hi - lo), (long long)got);
        return rc;
    });
}

int thr_group_detect_card(thr_group *g, const char *text, size_t len, int32_t final_chunk, int64_t max_blocks,
                          double *timestamps, int64_t *block_idx, thr_record *out, int64_t *n_blocks, int64_t *consumed) {
    if (!g || !text || !timestamps || !block_idx || !out || !n_blocks || !consumed || max_blocks < 0) return THR_ERR_INVALID;
    const int G = (int)g->w.size();
    const int64_t N = g->cfg.block_len, NT = g->cfg.n_templates;
    *n_blocks = 0;
    *consumed = 0;
    // one pass over the line headers fixes where every data line starts; the stripes are cut at line boundaries
    std::vector<int64_t> off((size_t)max_blocks);
    int64_t found = 0, used = 0, bad_line = -1;
    const int rc0 = card_scan_lines(text, len, (int32_t)N, final_chunk, max_blocks, timestamps, block_idx, off.data(), &found,
                                    &used, &bad_line, nullptr);
    if (rc0 != THR_OK) {
        g->err = ".card data line " + std::to_string(bad_line) + " is malformed";
        return rc0;
    }
    *consumed = used;
    if (found == 0) return THR_OK;
    const int64_t want = ((2 * N + 2) / 3) * 4;
    const int rc = group_run(g, [&](int i, thr_detector *d) -> int {
        int64_t lo, hi;
        stripe_of(found, G, i, &lo, &hi);
        if (hi <= lo) return THR_OK;
        // text range of the stripe: from the start of line lo's header (searched backwards from its payload) to the end of
        // line hi-1's payload; the per-stripe call re-parses its own headers
        const char *p0 = text + off[(size_t)lo];
        while (p0 > text && p0[-1] != '\n') --p0;
        const char *p1 = text + off[(size_t)(hi - 1)] + want;
        int64_t nb = 0, cons = 0;
        const int r = thr_detect_card(d, p0, (size_t)(p1 - p0), 1, hi - lo, timestamps + lo, block_idx + lo,
                                      out + (size_t)lo * NT, &nb, &cons);
        if (r == THR_OK && nb != hi - lo) return fail(d, THR_ERR_INVALID, "stripe of %lld lines returned %lld", (long long)(hi - lo), (long long)nb);
        return r;
    });
    if (rc == THR_OK) *n_blocks = found;
    return rc;
}

// Page-locked host buffer of n_devices * bytes_per_device bytes whose g-th part is placed on the NUMA node of the group's
// g-th GPU (mbind before first touch, then cudaHostRegister); a plain page-locked allocation where the host shows one node.
void *thr_group_host_alloc(thr_group *g, size_t bytes_per_device) {
    if (!g || !bytes_per_device) return nullptr;
    const size_t G = g->w.size(), page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t part = (bytes_per_device + page - 1) / page * page, total = part * G;
    void *base = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return nullptr;
    for (size_t i = 0; i < G; ++i) {
        const int node = g->w[i]->numa;
        if (node >= 0 && node < 64) {
            unsigned long mask = 1ul << node;
            syscall(SYS_mbind, (char *)base + i * part, part, 2 /* MPOL_BIND */, &mask, 65ul, 0u);   // best effort
        }
        std::memset((char *)base + i * part, 0, part);                                              // first touch
    }
    if (cudaHostRegister(base, total, cudaHostRegisterPortable) != cudaSuccess) {
        munmap(base, total);
        return nullptr;
    }
    return base;
}
void thr_group_host_free(thr_group *g, void *p, size_t bytes_per_device) {
    if (!g || !p) return;
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t part = (bytes_per_device + page - 1) / page * page;
    cudaHostUnregister(p);
    munmap(p, part * g->w.size());
}

// ---- .toad text (thrifty/toads_data.py:47-61) ------------------------------------------------------------------
// One line per detected block: "{rxid} {t:.6f} {block} {soa:.8f} {sample} {offset} {energy} {noise} {bin} {coffset}
// {cenergy} {cnoise}".  The fields without a format spec are what Python's str() prints for the values the Python layer
// holds (thrifty_b200/detect.py records_to_results): float32 record fields widened to Python floats print as the shortest
// string that round-trips the DOUBLE, the carrier energy / noise stay numpy.float32 (as in the reference) and print as the
// shortest string that round-trips the FLOAT.  Same digits as CPython / numpy (all are shortest-round-trip), same layout
// rule: fixed notation for 1e-4 <= |x| < 1e16 (numpy float32: < 1e6) with ".0" appended to integers, else d.ddde+XX.
}  // extern "C"
namespace {

template <class F>
char *py_float_str(char *out, F v) {
    if (v != v) { std::memcpy(out, "nan", 3); return out + 3; }
    if (std::isinf(v)) {
        if (v < 0) *out++ = '-';
        std::memcpy(out, "inf", 3);
        return out + 3;
    }
    char sci[48];
    const auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);   // shortest: d[.ddd]e[+-]XX
    char *p = sci;
    if (*p == '-') *out++ = *p++;
    char digits[32];
    int nd = 0;
    digits[nd++] = *p++;
    if (*p == '.') {
        ++p;
        while (*p != 'e') digits[nd++] = *p++;
    }
    ++p;                                           // 'e'
    int e = 0;
    std::from_chars(*p == '+' ? p + 1 : p, r.ptr, e);
    if (nd == 1 && digits[0] == '0') { std::memcpy(out, "0.0", 3); return out + 3; }
    const int decpt = e + 1;                       // position of the decimal point relative to the digit string
    // exponent notation: CPython's repr of a float switches at 1e16, numpy's str of a float32 scalar at 1e6
    // (CPython decides on the decimal exponent of the shortest digits, numpy on the value itself)
    const double av = std::fabs((double)v);
    const bool sci_notation = sizeof(F) == 4 ? (av >= 1e6 || av < 1e-4) : (decpt > 16 || decpt <= -4);
    if (sci_notation) {
        *out++ = digits[0];
        if (nd > 1) {
            *out++ = '.';
            std::memcpy(out, digits + 1, (size_t)nd - 1);
            out += nd - 1;
        }
        *out++ = 'e';
        *out++ = e < 0 ? '-' : '+';
        const int ae = e < 0 ? -e : e;
        if (ae < 10) *out++ = '0';
        out = std::to_chars(out, out + 8, ae).ptr;
        return out;
    }
    if (decpt <= 0) {                              // 0.000ddd
        *out++ = '0';
        *out++ = '.';
        for (int i = 0; i < -decpt; ++i) *out++ = '0';
        std::memcpy(out, digits, (size_t)nd);
        return out + nd;
    }
    if (decpt >= nd) {                             // ddd000.0
        std::memcpy(out, digits, (size_t)nd);
        out += nd;
        for (int i = nd; i < decpt; ++i) *out++ = '0';
        *out++ = '.';
        *out++ = '0';
        return out;
    }
    std::memcpy(out, digits, (size_t)decpt);       // dd.ddd
    out += decpt;
    *out++ = '.';
    std::memcpy(out, digits + decpt, (size_t)(nd - decpt));
    return out + (nd - decpt);
}

constexpr size_t TOAD_LINE_MAX = 1024;             // 10 fields of <= 26 characters + "%.6f" / "%.8f" of any double (<= 320 each)

char *toad_line(char *o, const thr_record &r, double ts, int32_t rxid, int32_t txid) {
    o = std::to_chars(o, o + 12, rxid).ptr;
    *o++ = ' ';
    if (txid != INT32_MIN) {                       // .toads: transmitter id after the receiver id (toads_data.py:57-60)
        o = std::to_chars(o, o + 12, txid).ptr;
        *o++ = ' ';
    }
    o = std::to_chars(o, o + 330, ts, std::chars_format::fixed, 6).ptr;          // "%.6f" (exact, like printf)
    *o++ = ' ';
    o = std::to_chars(o, o + 24, (long long)r.block_idx).ptr;
    *o++ = ' ';
    o = std::to_chars(o, o + 330, r.soa, std::chars_format::fixed, 8).ptr;       // "%.8f"
    *o++ = ' ';
    o = std::to_chars(o, o + 12, r.corr_sample).ptr;
    *o++ = ' ';
    o = py_float_str(o, (double)r.corr_offset);
    *o++ = ' ';
    o = py_float_str(o, (double)r.corr_energy);
    *o++ = ' ';
    o = py_float_str(o, (double)r.corr_noise);
    *o++ = ' ';
    o = std::to_chars(o, o + 12, r.carrier_bin).ptr;
    *o++ = ' ';
    o = py_float_str(o, (double)r.carrier_offset);
    *o++ = ' ';
    o = py_float_str(o, r.carrier_energy);         // numpy.float32 in the Python layer
    *o++ = ' ';
    o = py_float_str(o, r.carrier_noise);
    *o++ = '\n';
    return o;
}

}  // namespace
extern "C" {

// Formats the detected records (flags & THR_FLAG_CORR_DETECTED) of recs[0 .. n) (stride = records per block, template 0)
// as .toad lines into buf; *used = bytes written.  txids: NULL for a .toad, else one transmitter id per record (.toads).
// THR_ERR_NOMEM (and *used = bytes needed at most) if cap < n_detected * 1024.  Several threads for large n.
int thr_format_toad(const thr_record *recs, const double *timestamps, int64_t n, int64_t stride, int32_t rxid,
                    const int32_t *txids, char *buf, size_t cap, size_t *used) {
    if (!recs || !timestamps || !used || n < 0 || stride < 1) return THR_ERR_INVALID;
    int64_t n_det = 0;
    for (int64_t i = 0; i < n; ++i) n_det += (recs[i * stride].flags & THR_FLAG_CORR_DETECTED) ? 1 : 0;
    *used = (size_t)n_det * TOAD_LINE_MAX;
    if (!buf || cap < *used) return THR_ERR_NOMEM;
    int parts = (int)(n / 512);                    // ~0.5 ms of formatting per part
    parts = parts < 1 ? 1 : (parts > 8 ? 8 : parts);
    std::vector<char *> begin((size_t)parts), end((size_t)parts);
    std::vector<std::thread> th;
    int64_t det_before = 0;
    for (int p = 0; p < parts; ++p) {
        const int64_t lo = n * p / parts, hi = n * (p + 1) / parts;
        char *o0 = buf + (size_t)det_before * TOAD_LINE_MAX;
        for (int64_t i = lo; i < hi; ++i) det_before += (recs[i * stride].flags & THR_FLAG_CORR_DETECTED) ? 1 : 0;
        begin[(size_t)p] = o0;
        auto work = [=, &end] {
            char *o = o0;
            for (int64_t i = lo; i < hi; ++i) {
                const thr_record &r = recs[i * stride];
                if (r.flags & THR_FLAG_CORR_DETECTED) o = toad_line(o, r, timestamps[i], rxid, txids ? txids[i] : INT32_MIN);
            }
            end[(size_t)p] = o;
        };
        if (p + 1 < parts) th.emplace_back(work); else work();
    }
    for (auto &t : th) t.join();
    char *o = buf;
    for (int p = 0; p < parts; ++p) {              // close the gaps between the parts
        const size_t len = (size_t)(end[(size_t)p] - begin[(size_t)p]);
        if (begin[(size_t)p] != o) std::memmove(o, begin[(size_t)p], len);
        o += len;
    }
    *used = (size_t)(o - buf);
    return THR_OK;
}

void *thr_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void thr_host_free(void *p) {
    if (p) cudaFreeHost(p);
}
void *thr_device_alloc(int device, size_t bytes) {
    void *p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
void thr_device_free(int device, void *p) {
    if (!p) return;
    cudaSetDevice(device);
    cudaFree(p);
}
int thr_memcpy_h2d(int device, void *dst, const void *src, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return THR_ERR_CUDA;
    return cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? THR_OK : THR_ERR_CUDA;
}
int thr_memcpy_d2h(int device, void *dst, const void *src, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess) return THR_ERR_CUDA;
    return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? THR_OK : THR_ERR_CUDA;
}

}  // extern "C"
