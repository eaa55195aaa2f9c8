// detect_kernel.cuh -- fused per-block detect kernel for sm_100a.
//
// One persistent CTA per SM (N=16384) walks the batch; for each block it runs the whole
// Thrifty detect chain (thrifty/detect.py:60-78) without touching HBM in between:
//
//   raw u8 tile (TMA bulk copy -> smem, double buffered)
//   -> rawconv (block_data.py:38-52) -> FFT#1 -> |X|^2 sum + windowed arg-max
//   -> threshold (carrier_detect.py:61-115) -> Dirichlet LM fit (carrier_sync.py:150-196)
//   -> mix (carrier_sync.py:222-238) from the still-resident raw tile -> FFT#2
//   -> x conj(T)/N -> IFFT (soa_estimator.py:97-102) -> |c|^2 windowed arg-max (:137-143)
//   -> noise / threshold (:108-134) -> Gaussian interpolation (:159-170) -> 64-byte record.
//
// FFT: N = 32 * R2 * R3 decimation-in-frequency, in place in a per-CTA complex buffer
// (shared memory for N <= 16384, rows padded 16->17 so that every pass is bank-conflict free;
// an L2-resident global scratch for N = 32768).  Each pass is a radix-32/R2/R3 DFT held
// entirely in registers.  The forward transform leaves the spectrum digit-reversed across
// threads; the template spectrum is stored pre-permuted to match and the inverse transform
// runs the mirrored passes, so no reordering pass exists.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/thrifty_b200.h"
#include "dirichlet_lm.cuh"

// experiment switches (compile-time; defaults are the measured-best configuration)
#ifndef THR_WL23
#define THR_WL23 1          // warp-local hand-over between passes 2 and 3 (no CTA barrier)
#endif
#ifndef THR_TW3_MULTI
#define THR_TW3_MULTI 0     // the same in the multi-template kernels: measured -6.5 % (the inverse-side chain is paid once per
                            // template, the table look-up it replaces in the forward pass 2 only once per block)
#endif
#ifndef THR_SERVICE_T256
#define THR_SERVICE_T256 1      // (+12 % at N = 8192) service warpgroup also for the 2-CTAs-per-SM kernel
#endif
#ifndef THR_SERVICE_T128
#define THR_SERVICE_T128 1      // (+10 % at N = 4096) the same for the 4-CTAs-per-SM kernel (N = 4096): 64 x 256 = 128 x 96 + 128 x 32
#endif
#ifndef THR_T128_CTAS
#define THR_T128_CTAS 4         // resident CTAs per SM of that kernel (3: 80 x 256 = 128 x 128 + 128 x 32, no spills, but -4 %)
#endif
#ifndef THR_FIT_DEPTH_BIG
#define THR_FIT_DEPTH_BIG 4     // blocks of look-ahead of stage A over stage B where shared memory allows a 5-slot ring (T >= 256)
#endif
#ifndef THR_ROLL_ITEMS
#define THR_ROLL_ITEMS 1        // kernels with several CTAs per SM (N <= 8192) keep the loops over a thread's pass-2 / pass-3
                                // items rolled: those kernels are bound by instruction fetch (four unsynchronised CTAs walk
                                // the same straight-line code), and 2-4 copies of a radix-8/16 transform are the bulk of it
#endif
#ifndef THR_TW3
#define THR_TW3 1           // (+2 %) inter-pass twiddles W_M^{n3 k2} of FFT#2 / IFFT applied on the pass-3 side from a
                            // per-item register chain instead of the shared-memory table on the pass-2 side
#endif

namespace thr {

struct DetectParams {
    const uint8_t *raw;        // u8 interleaved I,Q: block b starts at raw + b*raw_stride (or nullptr)
    int64_t raw_stride;        // bytes between block starts: 2N (.card blocks) or 2(N-H) (contiguous stream)
    const float2  *iq;         // [n_blocks][N] complex64 (used when raw == nullptr)
    const int64_t *block_idx;  // [n_blocks] or nullptr (-> 0,1,2..)
    thr_record    *out;        // [n_blocks][n_templates]
    int n_blocks;
    int n_templates;
    const float2 *tpl_spec;    // [n_templates][N] conj(FFT(template))/N in kernel order
    const float  *tpl_energy;  // [n_templates] sum(template^2)
    float2 *scratch;           // per-CTA global scratch: [grid][N] FFT buffer (GMEM variant)
    float2 *xsave;             // per-CTA save area for X' when n_templates > 1: [grid][N]
    int win_start, win_len;    // carrier window: start index in [0,N), number of bins
    float c_const, c_snr, c_std;   // carrier threshold coefficients
    float k_const, k_snr, k_std;   // correlation threshold coefficients
    int corr_start, corr_stop, corr_len;
    int new_len;               // N - H
    int zoom;                  // 1: carrier window (+-3 bins) spans <= 128 bins and no stddev term -> pruned FFT#1
    int zoom_base;             // first bin b0 of the 128-bin zoom band: the block is pre-shifted by -b0 bins (0: none)
    int zoom_w0;               // window start inside the band: (win_start - zoom_base) mod N
    double fit_W;              // carrier_len (the Dirichlet fit runs in float64, dirichlet_lm.cuh)
    double fit_piW;            // pi * W, rounded as the reference's np.pi*W (carrier_sync.py:129)
    double fit_N;              // block_len
    // fastdet-semantics kernels (FASTDET = true) only:
    const float2 *tpl_shift;   // [win_len][N]: conj(T)[(k - kpeak) mod N]/N per carrier bin of the window, kernel
                               // order (the integer roll of fastdet/corr_detector.cpp:13-17,179 folded into the
                               // template), or nullptr -> gather from tpl_nat
    const float2 *tpl_nat;     // [N] conj(FFT(template))/N in natural order
    float2 *dbg_shifted_fft;   // optional: shifted spectrum X' (FFT#2), natural order, block b at + b * dbg_sfft_stride
    float2 *dbg_corr;          // optional: correlation c[0 .. corr_len), block b at + b * dbg_corr_stride (template 0)
    float  *dbg_fft_mag;       // optional [N] (single-block debug launches)
    int64_t dbg_sfft_stride;   // elements between blocks (0: single-block debug launch; N: thr_sync_batch)
    int64_t dbg_corr_stride;   // (0, or corr_len: thr_soa_batch)
    const float2 *in_sfft;     // STAGES == 2 kernels: [n_blocks][N] shifted spectra, natural order (the input)
};

template <int LOG2N_, int T_, bool GMEM_, bool FASTDET_ = false>
struct Cfg {
    static constexpr int LOG2N = LOG2N_;
    static constexpr int N = 1 << LOG2N_;
    static constexpr int T = T_;
    static constexpr bool GMEM = GMEM_;
    static constexpr int M = N / 32;                 // size of the sub-transforms after pass 1
    static constexpr int R3 = GMEM_ ? 32 : 16;       // last-pass radix
    static constexpr int R2 = M / R3;                // middle-pass radix
    static constexpr int S = 32 * R2;                // bin stride of the last pass (k = kb + S*k3)
    static constexpr int I1 = M / T;                 // work items per thread, pass 1 (radix 32)
    static constexpr int I2 = 32 * R3 / T;           // pass 2 (radix R2)
    static constexpr int I3 = 32 * R2 / T;           // pass 3 (radix R3)
    static_assert(M % T == 0 && I1 >= 1, "bad thread count");
    static_assert((32 * R3) % T == 0 && I2 >= 1, "bad thread count");
    static_assert((32 * R2) % T == 0 && I3 >= 1, "bad thread count");
    static_assert(R2 >= 1 && R2 <= 32 && R3 <= 32, "unsupported size");
    static_assert(!GMEM_ || (R2 == 32 && R3 == 32 && (32 * R3) % T_ == 0), "GMEM variant: natural pass-3 order assumed");
    // Shared-memory FFT buffer: rows of 16 complex values padded to 17 (136 bytes), so that a
    // half-warp touches 16 different 8-byte bank pairs both when it walks along a row (passes
    // 1, 2) and when it walks across rows (pass 3), and every access is base + immediate.
    static constexpr int ROW_BYTES = GMEM_ ? R3 * 8 : 136;
    static constexpr size_t BUF_BYTES = GMEM_ ? 0 : (size_t)(N / 16) * 136;
    // A dedicated service warpgroup runs the serial fit / tail while the workers go on with the next block; registers
    // are re-split with setmaxnreg (T == 512: workers 112, service 32).  Without it (T <= 64, fastdet flow at T <= 256)
    // warp 0 runs the serial parts inline and the other CTAs of the SM hide them.
    // ONE_CTA: one CTA per SM (the buffer takes more than half of the shared memory, or T == 512).
    static constexpr bool ONE_CTA = (T >= 512) || (BUF_BYTES > 100 * 1024);
    // setmaxnreg moves registers inside the pool the CTA was launched with (launch registers x LAUNCH_THREADS):
    // T == 512: 96 x 640 = 512 x 112 + 128 x 32.  T == 256 with two CTAs per SM (N = 8192): 80 x 384 = 256 x 104 + 128 x 32.
    // Measured at N = 8192: 133 -> 149 Gsamples/s with the service warpgroup; the fastdet flow (no fit, a short tail) is
    // better off with the registers: 196 vs 180 Gsamples/s, so it keeps the inline service.
    static constexpr bool SERVICE = ONE_CTA || (THR_SERVICE_T256 != 0 && T == 256 && !FASTDET_)
                                            || (THR_SERVICE_T128 != 0 && T == 128 && !FASTDET_);
    static constexpr int LAUNCH_THREADS = SERVICE ? T + 128 : T;
    static constexpr int MIN_CTAS = ONE_CTA ? 1 : (T >= 256 ? 2 : (T >= 128 ? (SERVICE ? THR_T128_CTAS : 4) : 8));
    // registers per worker after setmaxnreg: what the 64 K file leaves beside the 128 x 32 of the service warpgroup
    static constexpr int WORKER_REGS = !ONE_CTA ? (T == 256 ? 104 : (THR_T128_CTAS == 3 ? 128 : 96))
                                                : (T >= 512 ? 112 : (T >= 256 ? 232 : 240));
    static_assert(!SERVICE || T * WORKER_REGS + 128 * 32 <= (65536 / MIN_CTAS / LAUNCH_THREADS / 8 * 8) * LAUNCH_THREADS,
                  "setmaxnreg targets exceed the CTA's register pool (the kernel would dead-lock)");
    static constexpr int MAX_TPL = 32;               // templates per detector (tail mailbox size)
    // Stage A (FFT#1, carrier decision) runs FIT_DEPTH blocks ahead of stage B (mix, FFT#2, correlation): the float64
    // Levenberg-Marquardt fit between them (dirichlet_lm.cuh, ~6 K dependent instructions) takes about two block
    // periods on one warp, so the four warps of the service warpgroup take the blocks in turn: warp w fits the blocks
    // i = w (mod 4) and, after each fit, writes the records of block i - 2 (whose stage B finished a period earlier).
    // Stage B re-fetches its raw tile (L2-resident) instead of keeping it since stage A.
    static constexpr int FIT_DEPTH = (SERVICE && !FASTDET_) ? (THR_FIT_DEPTH_BIG != 0 && T >= 256 ? THR_FIT_DEPTH_BIG : 3) : 1;
    static constexpr int NSLOT = FIT_DEPTH + 1;      // FitSlot ring
    static constexpr int NFITW = (SERVICE && !FASTDET_) ? 4 : 1;   // warps running fits
    static_assert(NFITW == 1 || FIT_DEPTH <= NFITW, "a fit warp takes every NFITW-th block");
    // Passes 2 and 3 both work inside one k1 slab (M consecutive elements).  When every warp owns the
    // same slabs in both passes the hand-over 2 -> 3 (and 3' -> 2') only needs __syncwarp(): the warps
    // of a CTA then run the whole stretch  pass 2 -> 3 -> x conj(T) -> 3' -> 2'  without a CTA barrier
    // and drift apart, so the LDS/STS phases of one warp overlap the FMA phases of another.
    // Pass-2 items are (k1, n3) = w >> log2(R3), w & (R3-1) with w = tid + T*it, so a warp owns the slabs
    // k1 = 2*warp + h + (T/16)*it2 (R3 == 16; h = lane >> 4) or k1 = warp + (T/32)*it2 (R3 == 32); p3_item
    // hands the R2 rows of exactly those slabs to the same warp.  The template spectrum is stored in the
    // matching order (thr_create uses the same function).
    static constexpr bool WL23 = THR_WL23 != 0;
    // pass-3 item (row of R3 elements: k1*R2 + k2) of thread `tid`, iteration `it`
    __host__ __device__ static constexpr int p3_item(int tid, int it) {
        if (!WL23 || R3 != 16) return tid + T_ * it;     // R3 == 32: R2 == 32, the natural order matches
        const int r = it * 32 + (tid & 31);              // row number inside this warp's share
        const int sidx = r / R2, k2 = r % R2;            // slab number inside the share, row inside the slab
        const int k1 = 2 * (tid >> 5) + (sidx & 1) + (T_ / 16) * (sidx >> 1);
        return k1 * R2 + k2;
    }
    // pruned ("zoom") FFT#1, pass 3: with T == 512 and R2 == 32 a warp finds the 8 bins of its own slabs
    static constexpr bool ZOOM_WARPLOCAL = WL23 && (T_ == 512 && R2 == 32 && R3 == 16);
    static constexpr bool ZOOM_OK = (R2 >= 8 && R2 % 4 == 0);      // bins < 128 <=> k2 < 4, k3 == 0
    static constexpr size_t smem_bytes(bool multi) { // must cover the carve-up in detect_kernel
        return BUF_BYTES + 2 * (size_t)(2 * N) + (size_t)M * 8
               + NSLOT * 320 + 2 * 32 + 2 * (multi ? MAX_TPL : 1) * 32 + 256 + 512 + 256 + 64 + NFITW * sizeof(lm::Rows);
    }
};

// ------------------------------------------------------------------ small helpers
// Packed FP32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): a complex number is one aligned
// register pair; operand swizzles (LO_HI), per-half negation and scalar broadcast are folded
// into the instruction by ptxas, so a complex add is 1 instruction and a complex multiply 2.
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 rot_mj(float2 v) { return make_float2(v.y, -v.x); }   // v * (-i)
__device__ __forceinline__ float2 rot_pj(float2 v) { return make_float2(-v.y, v.x); }   // v * (+i)
__device__ __forceinline__ float2 f2scale(float2 v, float h) { return __fmul2_rn(v, make_float2(h, h)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {    // a * w
    return __ffma2_rn(rot_pj(a), make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 w) {   // a * conj(w)
    return __ffma2_rn(rot_mj(a), make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}
// v * (c - i s) (forward twiddle) or v * (c + i s) (INV)
template <bool INV>
__device__ __forceinline__ float2 mul_tw(float2 v, float c, float s) {
    return __ffma2_rn(INV ? rot_pj(v) : rot_mj(v), make_float2(s, s), __fmul2_rn(v, make_float2(c, c)));
}
// cis(pi * x): cos(pi x) + i sin(pi x), exact range reduction
__device__ __forceinline__ float2 cispi(float x) {
    float s, c;
    sincospif(x, &s, &c);
    return make_float2(c, s);
}
// Read-only global load that does not allocate in L1: the template spectrum (128 KB per block, re-read from L2 by every
// block) would otherwise flush the few KB of L1 left beside the shared-memory carve-out, which the service warp's
// float64 fit needs for its register spills.
__device__ __forceinline__ float2 ldg_stream(const float2 *ptr) {
    float2 v;
#ifdef THR_EXP_LDG
    v = __ldg(ptr);
#else
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(ptr));
#endif
    return v;
}
__host__ __device__ constexpr int brev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

// cos / sin of 2*pi*q/32, q in [0,16).  Flat switches (no recursion) so that the calls always
// inline and fold to immediates once the butterfly loops are unrolled.
__host__ __device__ __forceinline__ constexpr float cos32(int q) {
    switch (q) {
        case 0: return 1.0f;
        case 1: return 0.98078528040323043f;
        case 2: return 0.92387953251128674f;
        case 3: return 0.83146961230254524f;
        case 4: return 0.70710678118654757f;
        case 5: return 0.55557023301960229f;
        case 6: return 0.38268343236508984f;
        case 7: return 0.19509032201612833f;
        case 8: return 0.0f;
        case 9: return -0.19509032201612819f;
        case 10: return -0.38268343236508973f;
        case 11: return -0.55557023301960196f;
        case 12: return -0.70710678118654746f;
        case 13: return -0.83146961230254535f;
        case 14: return -0.92387953251128674f;
        case 15: return -0.98078528040323043f;
        default: return 0.0f;
    }
}
__host__ __device__ __forceinline__ constexpr float sin32(int q) {
    switch (q) {
        case 0: return 0.0f;
        case 1: return 0.19509032201612825f;
        case 2: return 0.38268343236508978f;
        case 3: return 0.55557023301960218f;
        case 4: return 0.70710678118654746f;
        case 5: return 0.83146961230254524f;
        case 6: return 0.92387953251128674f;
        case 7: return 0.98078528040323043f;
        case 8: return 1.0f;
        case 9: return 0.98078528040323043f;
        case 10: return 0.92387953251128674f;
        case 11: return 0.83146961230254546f;
        case 12: return 0.70710678118654757f;
        case 13: return 0.55557023301960218f;
        case 14: return 0.38268343236508989f;
        case 15: return 0.19509032201612861f;
        default: return 0.0f;
    }
}

// acc + v * W_32^q (forward, q in [0,32)): 2 packed FMAs, the rotation is an operand swizzle
template <int Q>
__device__ __forceinline__ float2 fma_tw32(float2 acc, float2 v) {
    constexpr int q = Q & 31;
    if constexpr (q == 0) return f2add(acc, v);
    else if constexpr (q == 8) return f2add(acc, rot_mj(v));
    else if constexpr (q == 16) return f2sub(acc, v);
    else if constexpr (q == 24) return f2add(acc, rot_pj(v));
    else {
        constexpr float sg = q >= 16 ? -1.0f : 1.0f;           // W^(q) = -W^(q-16)
        constexpr float c = sg * cos32(q & 15), sn = sg * sin32(q & 15);
        return __ffma2_rn(rot_mj(v), make_float2(sn, sn), __ffma2_rn(v, make_float2(c, c), acc));
    }
}

// In-register radix-2 decimation-in-time FFT of size R (power of two <= 32) on packed complex
// values.  Input n must be placed at index brev(n); output k comes out at index k.
// INV = false: forward (e^{-i...}); INV = true: inverse (e^{+i...}, unnormalised).
// Butterfly (a, b, w) -> (a + w b, a - w b) costs 3 packed instructions when w is non-trivial:
//   u = fma(rot(b), s, fma(b, c, a));  v = fma(a, 2, -u)
// and 2 when w is 1 or -+i (the rotation is an operand swizzle).
template <int R, bool INV>
__device__ __forceinline__ void fft_dit(float2 (&x)[R]) {
#pragma unroll
    for (int len = 2; len <= R; len <<= 1) {
        const int half = len >> 1;
#pragma unroll
        for (int base = 0; base < R; base += len) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const int ia = base + i, ib = ia + half;
                const int q = i * (32 / len);            // twiddle W_len^i = W_32^q, q in [0,16)
                const float2 a = x[ia], b = x[ib];
                if (q == 0) {
                    x[ia] = f2add(a, b);
                    x[ib] = f2sub(a, b);
                } else if (q == 8) {                     // w = -i (forward) / +i (inverse)
                    const float2 t = INV ? rot_pj(b) : rot_mj(b);
                    x[ia] = f2add(a, t);
                    x[ib] = f2sub(a, t);
                } else {
                    const float c = cos32(q), sn = sin32(q);       // w = c -+ i sn
                    const float2 u = __ffma2_rn(INV ? rot_pj(b) : rot_mj(b), make_float2(sn, sn),
                                                __ffma2_rn(b, make_float2(c, c), a));
                    x[ia] = u;
                    x[ib] = __ffma2_rn(a, make_float2(2.0f, 2.0f), make_float2(-u.x, -u.y));
                }
            }
        }
    }
}

// ------------------------------------------------------------------ mbarrier / TMA bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                             uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ block-wide reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}
// pack (non-negative float value, index key): larger value wins, smaller key wins ties
__device__ __forceinline__ unsigned long long pack_cand(float v, uint32_t key) {
    return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xffffffffu - key);
}

struct RedOut {
    float s0, s1;               // two running sums
    unsigned long long best;    // packed arg-max candidate
};

// All threads get the block totals.  `red` is >= (T/32)*4 words of smem scratch.  Contains
// two __syncthreads(); the caller guarantees `red` is not in use by a previous reduction.
template <int T>
__device__ __forceinline__ RedOut block_reduce(float s0, float s1, unsigned long long best, uint32_t *red,
                                               int tid) {
    constexpr int NW = T / 32;
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    best = warp_max_u64(best);
    const int w = tid >> 5;
    if ((tid & 31) == 0) {
        red[w * 4 + 0] = __float_as_uint(s0);
        red[w * 4 + 1] = __float_as_uint(s1);
        red[w * 4 + 2] = (uint32_t)(best >> 32);
        red[w * 4 + 3] = (uint32_t)best;
    }
    __syncthreads();
    RedOut r;
    r.s0 = 0.f;
    r.s1 = 0.f;
    r.best = 0ull;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        r.s0 += __uint_as_float(red[i * 4 + 0]);
        r.s1 += __uint_as_float(red[i * 4 + 1]);
        unsigned long long b = ((unsigned long long)red[i * 4 + 2] << 32) | red[i * 4 + 3];
        r.best = b > r.best ? b : r.best;
    }
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------ Dirichlet-kernel fit
// Least-squares fit of A*|D(x - d)| to the 7 magnitudes at x = -3..3 (carrier_sync.py:150-196: scipy curve_fit
// 'lm' from p0 = (y[3], 0)) = MINPACK lmdif in float64, restated in dirichlet_lm.cuh.  Executed by one warp: every
// lane runs the (scalar, warp-uniform) iteration; the 7 rows of the problem live in shared memory (lm::Rows), lane
// r < 7 does the element-wise work of row r, lanes 8..14 evaluate the sines of the look-ahead offset.
// y: magnitude of point (lane & 7).  Returns the offset.
// lm::gn_fit for one warp: lane r (mod 8) < 7 owns point r, all four groups of 8 lanes compute the same thing.
template <class T>
struct GnWarp {
    T A, d;
    double slack;
    unsigned pattern;
    bool ok;
};
template <class T>
__device__ __forceinline__ T xor_sum8(T v) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <class T>
__device__ __forceinline__ GnWarp<T> gn_fit_warp(T y, int lane, T piW, T N, T W, T a0, T d0, T step_tol, int settle,
                                                  unsigned pattern_in, int max_iter) {
    const bool active = (lane & 7) < 7;
    const T xi = (T)(min(lane & 7, 6) - 3);
    T A = a0, d = d0;
    GnWarp<T> q;
    q.ok = false;
    q.slack = 1.0;
    q.pattern = pattern_in;
    T cost_prev = (T)3e38;
    T saa = 0, sdd = 0, sad = 0, cost = 0;
    bool converged = false;
#pragma unroll 1
    for (int it = 0; it < max_iter; ++it) {
        T D, Dz;
        lm::kernel_deriv<T>(xi - d, piW, N, W, D, Dz);
        const T g = active ? fabs(D) : (T)0, jd = active ? A * (D < (T)0 ? Dz : -Dz) : (T)0, r = active ? A * g - y : (T)0;
        saa = xor_sum8(g * g);
        sad = xor_sum8(g * jd);
        sdd = xor_sum8(jd * jd);
        const T sar = xor_sum8(g * r), sdr = xor_sum8(jd * r);
        cost = xor_sum8(r * r);
        const unsigned pattern = __ballot_sync(0xffffffffu, active && D < (T)0) & 0x7fu;
        // a point on a null of the kernel at lmdif's starting guess: not vouched for (see lm::gn_fit)
        if (it == 0 && settle == 1 && __any_sync(0xffffffffu, active && (lane & 7) != 3 && g < (T)lm::START_NULL)) break;
        if (it == settle) q.pattern = pattern;
        const T slop = sizeof(T) == 4 ? (T)1.00002 : (T)1.00000000001;
        if (!(cost <= cost_prev * slop) || (it > settle - (settle == 0) && pattern != q.pattern)) break;
        cost_prev = cost;
        const T det = saa * sdd - sad * sad;
        if (!(det > (T)0)) break;
        const T dA = -(sdd * sar - sad * sdr) / det, dd = -(saa * sdr - sad * sar) / det;
        A += dA;
        d += dd;
        if (it >= settle && fabs(dd) < step_tol && fabs(dA) <= step_tol * fabs(A)) {
            converged = true;
            break;
        }
    }
    q.A = A;
    q.d = d;
    if (converged) {
        const double det = (double)saa * (double)sdd - (double)sad * (double)sad;
        q.slack = sqrt(2.0 * lm::TOL * (double)cost * ((double)saa / det));
        q.ok = true;
    }
    return q;
}

struct WarpExec {                // lm::fit on one warp: lane r < 7 owns row r of the shared lm::Rows
    int lane;
    template <class F>
    __device__ __forceinline__ void each(int from, F &&f) const {
        if (lane < lm::M && lane >= from) f(lane);
    }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};
__device__ __forceinline__ float dirichlet_fit(float y, int lane, const DetectParams &p, lm::Rows &w) {
    const int li = min(lane & 7, 6);
    const double xi = (double)(li - 3);
    const double piW = p.fit_piW, N = p.fit_N, W = p.fit_W;
    // lanes 0..7 evaluate the first offset of a call, lanes 8..15 the second (slot 7 is a dummy)
    auto weights = [&](double da, double db, double *ga, double *gb) {
        if (lane < 16) {
            const double g = lm::weight(xi - (lane < 8 ? da : db), piW, N, W);
            (lane < 8 ? ga : gb)[lane & 7] = g;
        }
        __syncwarp();
    };
#ifndef THR_EXP_NOQUICK
    // short cut (dirichlet_lm.cuh, lm::quick_fit): Gauss-Newton to the least-squares minimum (float steps first, double
    // to finish), accepted where lmdif provably stops within 3e-5 bins of it (every block of a well-conditioned geometry
    // such as the example configuration).  Same iteration as lm::gn_fit, laid out for a warp: one lane per point, the
    // six sums of the normal equations by xor-shuffles, nothing in shared memory.
    if (N / W <= lm::QUICK_MAX_LOBE) {
        const GnWarp<float> p1 = gn_fit_warp<float>(y, lane, (float)piW, (float)N, (float)W, __shfl_sync(0xffffffffu, y, 3), 0.f,
                                                    lm::QUICK_STEP_F32, 1, 0u, lm::QUICK_MAXIT);
#ifdef THR_EXP_F32ONLY          // timing experiment: what does the double-precision finish cost?
        if (p1.ok) return p1.d;
#endif
        if (p1.ok) {
            const GnWarp<double> p2 = gn_fit_warp<double>((double)y, lane, piW, N, W, (double)p1.A, (double)p1.d, lm::QUICK_STEP,
                                                          0, p1.pattern, 4);
            if (p2.ok && p2.slack < lm::QUICK_SLACK && fabs(p2.d) < 1.0) return (float)p2.d;
        }
    }
#endif
    if (lane < 8) w.y[lane] = (double)y;
    __syncwarp();
    const lm::Result res = lm::fit(WarpExec{lane}, weights, w, w.y[3], 0.0);
    __syncwarp();                                  // the rows may be overwritten by the next fit
    return (float)res.offset;
}

// ------------------------------------------------------------------ rawconv
// (b - 127.4f) / 128 exactly as numpy float32 does it (block_data.py:38-52), without I2F:
// 0x4B000000 | b is the float 2^23 + b; subtracting 2^23 is exact, and the fused multiply-add
// rounds the exact value (b - 127.4f) * 2^-7, which is representable.
__device__ __forceinline__ float2 rawconv(uint32_t w16) {
    constexpr float c = -127.4f * 0.0078125f;
    const float2 f = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(w16, 0x4B000000u, 0x7650)),
                                            __uint_as_float(__byte_perm(w16, 0x4B000000u, 0x7651))),
                                make_float2(-8388608.0f, -8388608.0f));
    return __ffma2_rn(f, make_float2(0.0078125f, 0.0078125f), make_float2(c, c));
}

// ------------------------------------------------------------------ named barriers
// The CTA has T "main" threads (the FFT workers) and one extra "service" warp.  Main threads
// synchronise among themselves on barrier BAR_MAIN; requests to / completions from the
// service warp use arrive/sync pairs on per-parity barrier ids, so neither side can run two
// generations ahead on the same id.
// Every id that a service warp waits on belongs to that warp alone: two warps waiting on one id at different
// generations could steal each other's generation.  ids: BAR_MAIN; BAR_TAILREQ / BAR_FITREQ / BAR_FITDONE + w with
// w = service warp (detect flow: block % 4 for the fit, (block + 2) % 4 for the records) or block & 1 (single
// service warp: FASTDET flow, detect2x_kernel).
enum : int { BAR_MAIN = 1, BAR_TAILREQ = 2, BAR_FITREQ = 6, BAR_FITDONE = 10 };
__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// per-parity mailbox between the main threads and the service warp
struct FitSlot {            // written by main after FFT#1 (A stage), completed by the service warp
    float mags[8];          // 7 magnitudes around the carrier peak
    int   kpeak;
    int   carrier;          // carrier detected?
    float peak_mag, noise_c, sig_energy1;
    float delta;            // <- service warp
    float pad[2];
    float2 rho[32];         // <- service warp: exp(-2 pi i (k+delta) n1 / 32)
};
struct TailSlot {           // written by main at the end of the correlation stage
    float peak_cp;          // |c|^2 at the peak
    int   s;                // peak lag
    float unused0;
    float pa, pc;           // |c|^2 at s-1, s+1
    float c1, c2;           // sum |c|, sum |c|^2 over [0, corr_len) (stddev threshold term only)
    float pad;
};
struct TailHdr {            // carrier fields copied for the record
    int   kpeak, carrier;
    float peak_mag, noise_c, sig_energy1, delta;
    float pad[2];
};

static_assert(sizeof(FitSlot) == 320 && sizeof(TailSlot) == 32 && sizeof(TailHdr) == 32, "mailbox layout");

// Block-wide arg-max (+ optional sums) over the T worker threads; two BAR_MAIN syncs.
//   vbits: bit pattern of this thread's best (non-negative) value, 0 if it has none
//   find_key(gbits): called only by threads whose vbits equals the block maximum; returns the
//                    smallest key (bin / lag) at which this thread holds that value
// Returns the maximum and the smallest key over all threads holding it (numpy's first-max rule).
struct ArgOut {
    uint32_t vbits, key;
    float s0, s1;
};
template <int T, bool SUMS, class F>
__device__ __forceinline__ ArgOut main_argmax(uint32_t vbits, float s0, float s1, uint32_t *red, int tid,
                                              F &&find_key) {
    constexpr int NW = T / 32;
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, vbits);
    if (SUMS) {
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
    }
    const int w = tid >> 5;
    if ((tid & 31) == 0) {
        red[w] = wmax;
        if (SUMS) {
            red[16 + w] = __float_as_uint(s0);
            red[32 + w] = __float_as_uint(s1);
        }
        if (tid == 0) red[48] = 0xffffffffu;
    }
    bar_sync(BAR_MAIN, T);
    ArgOut r;
    r.vbits = 0u;
    r.s0 = 0.f;
    r.s1 = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        r.vbits = max(r.vbits, red[i]);
        if (SUMS) {
            r.s0 += __uint_as_float(red[16 + i]);
            r.s1 += __uint_as_float(red[32 + i]);
        }
    }
    if (vbits == r.vbits) atomicMin(&red[48], find_key(r.vbits));
    bar_sync(BAR_MAIN, T);
    r.key = red[48];
    return r;
}

// ------------------------------------------------------------------ the kernel
// Launched with T + 32 threads: warps 0..T/32-1 are the FFT workers, the last warp is the
// service warp (Dirichlet fit + mix phasor table, scalar tail + record store).  Software
// pipeline per CTA:   A(b+1) || fit(b)   ->   B(b) || tail(b-1)
//   A(b): raw tile -> FFT#1 -> |X|^2, arg-max, carrier decision, 7 magnitudes posted
//   B(b): mix + FFT#2 -> x conj(T)/N -> IFFT -> |c|^2 arg-max, neighbours posted
// MULTI = false: exactly one template (no template loop, no X' save area).
// FASTDET = true: the semantics of the reference's native twin (fastcard + fastdet): decisions on powers,
// integer-bin carrier shift folded into the template spectrum (so FFT #2 disappears: 2 transforms per
// block), parabolic carrier offset, +-0.5 clip -- see the FASTDET section below.
// STAGES: 0 = the whole chain; 1 = stop at the stage boundary of the reference's Synchronizer (carrier_sync.py:52-76:
// carrier decision, fit, mix, FFT#2 -> shifted spectrum + carrier fields of the record; thr_sync_batch); 2 = start
// there (soa_estimator.py:78-92: shifted spectrum in -> correlation, peak, threshold, interpolation; thr_soa_batch).
template <int LOG2N, int T, bool GMEM, bool MULTI, bool FASTDET = false, int STAGES = 0>
__global__ void __launch_bounds__(Cfg<LOG2N, T, GMEM, FASTDET>::LAUNCH_THREADS, Cfg<LOG2N, T, GMEM, FASTDET>::MIN_CTAS)
detect_kernel(const __grid_constant__ DetectParams p) {
    static_assert(STAGES == 0 || !FASTDET, "stage-boundary kernels follow the Python path's semantics");
    using C = Cfg<LOG2N, T, GMEM, FASTDET>;
    constexpr bool SERVICE = C::SERVICE;
    constexpr bool TW3 = (THR_TW3 != 0) && (!MULTI || THR_TW3_MULTI != 0) && !FASTDET && C::R3 == 16 && C::R2 > 1;
    constexpr int N = C::N, M = C::M, R2 = C::R2, R3 = C::R3, S = C::S;
    constexpr int I1 = C::I1, I2 = C::I2, I3 = C::I3;
    constexpr bool ROLL = (THR_ROLL_ITEMS != 0) && !C::ONE_CTA;
    constexpr int UNROLL_I2 = ROLL ? 1 : I2;                   // #pragma unroll factors of the item loops
    constexpr int UNROLL_I3 = (ROLL && !FASTDET) ? 1 : I3;     // (FASTDET keeps its pass-3 outputs in registers per item)
    constexpr int LOG2M = ilog2(M), LOG2R3 = ilog2(R3), LOG2R2 = ilog2(R2), LOG2S = ilog2(S);
    constexpr uint32_t RAW_BYTES = 2u * N;
    constexpr int NTHREADS = T + 32;     // participants of the worker<->service barriers

    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;

    // ---- shared memory carve-up (Cfg::smem_bytes() must cover it)
    size_t off = 0;
    unsigned char *bufc;                 // FFT buffer, byte-addressed
    if (GMEM) {
        bufc = reinterpret_cast<unsigned char *>(p.scratch + (size_t)blockIdx.x * N);
    } else {
        bufc = smem;
        off += C::BUF_BYTES;
    }
    unsigned char *raw_s = smem + off;
    off += 2 * (size_t)RAW_BYTES;
    float2 *tw2 = reinterpret_cast<float2 *>(smem + off);
    off += (size_t)M * 8;
    FitSlot *fitslot = reinterpret_cast<FitSlot *>(smem + off);      // [NSLOT]
    off += C::NSLOT * sizeof(FitSlot);
    TailHdr *tailhdr = reinterpret_cast<TailHdr *>(smem + off);      // [2]
    off += 2 * sizeof(TailHdr);
    TailSlot *tailslot = reinterpret_cast<TailSlot *>(smem + off);   // [2][TPL_SLOTS]
    constexpr int TPL_SLOTS = MULTI ? C::MAX_TPL : 1;                // tail mailbox entries per block
    off += 2 * (size_t)TPL_SLOTS * sizeof(TailSlot);
    uint32_t *red = reinterpret_cast<uint32_t *>(smem + off);        // 64 words reduction scratch
    off += 256;
    float *zpow = reinterpret_cast<float *>(smem + off);             // |X[b0 + k]|^2, k < 128 (zoom path)
    off += 512;
    float2 *zrho = reinterpret_cast<float2 *>(smem + off);           // W_32^{b0 n1}: row phasors of the zoom pre-shift
    off += 256;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + off);       // 2 barriers
    off += 64;
    lm::Rows *fitrows = reinterpret_cast<lm::Rows *>(smem + off);    // [NFITW] row workspaces of the Dirichlet fit

    // Launch-invariant switches are re-read from the kernel parameters (constant bank, uniform
    // datapath) wherever they are used: held in registers they get spilled, and a spill reload
    // misses the small L1 that is left next to 207 KB of shared memory (long-scoreboard stalls).
#define use_raw (p.raw != nullptr)
#define need_std_c (p.c_std != 0.f)
#define need_std_k (p.k_std != 0.f)
    // blocks of this CTA: blockIdx.x + i * gridDim.x for i >= 0 while that is < n_blocks
    auto has_block = [&](int i) -> bool { return (int)blockIdx.x + i * (int)gridDim.x < p.n_blocks; };
    const int n_tpl = MULTI ? p.n_templates : 1;

    // Launches are independent of each other: let a dependent (next-batch) launch start as soon as
    // SMs drain (no-op unless the launch carries the programmatic-serialization attribute).
    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    // inter-pass twiddle table W_M^{n3 k2} (all threads help)
    for (int idx = tid; idx < M; idx += C::LAUNCH_THREADS) {
        const int k2 = idx >> LOG2R3, n3 = idx & (R3 - 1);
        const int e = (n3 * k2) & (M - 1);
        tw2[idx] = cispi(-2.0f * (float)e / (float)M);
    }
    if (tid < 32) zrho[tid] = cispi(-2.0f * (float)((p.zoom_base * tid) & 31) * 0.03125f);
    __syncthreads();

    // =====================================================================================
    // serial work: Dirichlet fit + mix phasor table, scalar tail + record store.  Executed by
    // the service warp (SERVICE) or inline by warp 0 of the workers.
    // =====================================================================================
    auto do_fit = [&](int q, lm::Rows &rows) {
        FitSlot &fs = fitslot[q];
        if (STAGES != 2 && fs.carrier) {
            const float y = (lane & 7) < 7 ? fs.mags[lane & 7] : 0.f;
#ifdef THR_EXP_NOFIT            // timing experiment only (wrong offsets): what the float64 fit costs the workers
            const float d = 0.f * y;
#else
            const float d = dirichlet_fit(y, lane, p, rows);
#endif
            // mix phasors of the 32 pass-1 rows: rho[n1] = exp(-2 pi i (k+d) n1 / 32)
            const int e = (fs.kpeak * lane) & 31;
            const float turns = -((float)e * 0.03125f) - d * ((float)lane * 0.03125f);
            fs.rho[lane] = cispi(2.f * turns);
            if (lane == 0) fs.delta = d;
        }
    };
    auto do_tail = [&](int i, int q) {
        const TailHdr &h = tailhdr[q];
        const int blk = (int)blockIdx.x + i * (int)gridDim.x;
        if (lane < n_tpl) {
            const int tpl = lane;
            const int64_t bidx = p.block_idx ? p.block_idx[blk] : (int64_t)blk;
            thr_record rec;
            rec.block_idx = bidx;
            rec.carrier_bin = h.kpeak;
            rec.carrier_energy = h.peak_mag;
            rec.carrier_noise = h.noise_c;
            rec.template_idx = tpl;
            rec.reserved = 0.f;
            if (!h.carrier || STAGES == 1) {      // no correlation ran for this block
                rec.soa = __longlong_as_double(0x7ff8000000000000ll);
                rec.carrier_offset = h.carrier ? h.delta : 0.f;
                rec.corr_sample = -1;
                rec.corr_offset = __int_as_float(0x7fc00000);
                rec.corr_energy = __int_as_float(0x7fc00000);
                rec.corr_noise = __int_as_float(0x7fc00000);
                rec.flags = h.carrier ? THR_FLAG_CARRIER_DETECTED : 0u;
                rec.signal_energy = h.sig_energy1;
            } else {
                // scalar tail (soa_estimator.py:78-134,159-170); float32 except the SoA itself
                const TailSlot &ts = tailslot[q * TPL_SLOTS + tpl];
                const float pa = ts.pa, pc = ts.pc;
                const float peak_mag_k = sqrtf(ts.peak_cp);
                // mean |X'|^2 (soa_estimator.py:111) == sum |x|^2 == mean |X|^2 of FFT#1: the mix is a
                // unit-modulus rotation and both FFTs are unitary up to N (Parseval)
                const float sig_energy = h.sig_energy1;
                const float noise_pw = (sig_energy * p.tpl_energy[tpl] - ts.peak_cp) / (float)N;
                const float noise_k = sqrtf(noise_pw);                 // NaN if negative
                float var_k = 0.f;
                if (need_std_k) {
                    const float mean = ts.c1 / (float)p.corr_len;
                    var_k = fmaxf(ts.c2 / (float)p.corr_len - mean * mean, 0.f);
                }
                const float thr_k = sqrtf(p.k_const + p.k_snr * (noise_k * noise_k) + p.k_std * var_k);
                const bool detected = peak_mag_k > thr_k;
                float offset = 0.f;
                if (detected && ts.s > 0 && ts.s < p.corr_len - 1) {
                    // a,b,c = ln|c|; offset = 0.5 (c-a) / (2b-a-c) with |c| = sqrt(power)
                    const float num = logf(pc / pa);
                    const float den = logf((ts.peak_cp / pa) * (ts.peak_cp / pc));
                    offset = fminf(fmaxf(0.5f * num / den, -0.6f), 0.6f);
                }
                rec.soa = (double)p.new_len * (double)bidx + (double)ts.s + (double)offset;
                rec.carrier_offset = h.delta;
                rec.corr_sample = ts.s;
                rec.corr_offset = offset;
                rec.corr_energy = peak_mag_k;
                rec.corr_noise = noise_k;
                rec.flags = (STAGES == 2 ? 0u : THR_FLAG_CARRIER_DETECTED) | (detected ? THR_FLAG_CORR_DETECTED : 0u);
                rec.signal_energy = sig_energy;
            }
            p.out[(size_t)blk * n_tpl + tpl] = rec;
        }
    };

    // scalar tail of the fastdet semantics (fastcard/cardet.c:22-40, fastdet/corr_detector.cpp:88-175,
    // fastdet/fastdet.cpp:184-206).  TailHdr carries POWERS here: peak_mag = max |X|^2, noise_c = noise power,
    // sig_energy1 = fft_sum / N, pad[0..1] = |X|^2 at the bins next to the peak.
    auto do_tail_fd = [&](int i, int q) {
        const TailHdr &h = tailhdr[q];
        const int blk = (int)blockIdx.x + i * (int)gridDim.x;
        if (lane == 0) {
            const int64_t bidx = p.block_idx ? p.block_idx[blk] : (int64_t)blk;
            thr_record rec;
            rec.block_idx = bidx;
            rec.carrier_bin = h.kpeak;
            rec.carrier_energy = sqrtf(h.peak_mag);                 // fastdet.cpp:204 prints sqrt(max)
            rec.carrier_noise = sqrtf(h.noise_c);
            rec.template_idx = 0;
            rec.reserved = 0.f;
            rec.signal_energy = h.sig_energy1;
            if (!h.carrier) {
                rec.soa = __longlong_as_double(0x7ff8000000000000ll);
                rec.carrier_offset = 0.f;
                rec.corr_sample = -1;
                rec.corr_offset = __int_as_float(0x7fc00000);
                rec.corr_energy = __int_as_float(0x7fc00000);
                rec.corr_noise = __int_as_float(0x7fc00000);
                rec.flags = 0u;
            } else {
                // corr_detector.cpp:88-101 parabolic interpolation on sqrt(power), clipped to +-0.5
                const float ca = sqrtf(h.pad[0]), cb = sqrtf(h.peak_mag), cc = sqrtf(h.pad[1]);
                const float coff = fminf(fmaxf((cc - ca) / (4.f * cb - 2.f * ca - 2.f * cc), -0.5f), 0.5f);
                const TailSlot &ts = tailslot[q * TPL_SLOTS];
                const float pa = ts.pa, pc = ts.pc;
                // corr_detector.cpp:118-125: the peak power arrives as size_t (truncated), noise clamped at 0
                float noise_pw = (h.sig_energy1 * p.tpl_energy[0] - truncf(ts.peak_cp)) / (float)N;
                noise_pw = noise_pw < 0.f ? 0.f : noise_pw;
                const float thr_k = p.k_const + p.k_snr * noise_pw;                      // :158
                const bool detected = ts.peak_cp > thr_k;
                float offset = 0.f;
                if (detected && ts.s > 0 && ts.s < p.corr_len - 1) {
                    // :103-116 Gaussian interpolation on ln sqrt(power), clipped to +-0.5
                    const float num = logf(pc / pa);
                    const float den = logf((ts.peak_cp / pa) * (ts.peak_cp / pc));
                    offset = fminf(fmaxf(0.5f * num / den, -0.5f), 0.5f);
                }
                rec.soa = (double)p.new_len * (double)bidx + (double)ts.s + (double)offset;   // fastdet.cpp:184
                rec.carrier_offset = coff;
                rec.corr_sample = ts.s;
                rec.corr_offset = offset;
                rec.corr_energy = sqrtf(ts.peak_cp);
                rec.corr_noise = sqrtf(noise_pw);
                rec.flags = THR_FLAG_CARRIER_DETECTED | (detected ? THR_FLAG_CORR_DETECTED : 0u);
            }
            p.out[blk] = rec;
        }
    };

    if constexpr (SERVICE) {
        if (tid >= T) {
            // service warpgroup: hand registers to the workers (FASTDET: only its first warp works)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
            const int sw = (tid - T) >> 5;                    // warp of the service warpgroup
            if (FASTDET && sw > 0) return;
            if constexpr (FASTDET) {
                for (int i = 0; has_block(i); ++i) {
                    const int q = i & 1;
                    bar_sync(BAR_TAILREQ + q, NTHREADS);                  // block i posted
                    do_tail_fd(i, q);
                    bar_arrive(BAR_FITDONE + q, NTHREADS);                // mailbox q may be reused
                }
                return;
            }
            if constexpr (!FASTDET) {
                // service warp sw: fit of block i = sw, sw + 4, ... (ring slot sw), then the records of block i - 2;
                // FITDONE(sw) tells the workers both that fit(i) is there and that tail mailbox i & 1 is free again
                for (int i = sw;; i += C::NFITW) {
                    const bool has_fit = has_block(i), has_tail = i >= 2 && has_block(i - 2);
                    if (!has_fit && !has_tail) break;
                    if (has_fit) {
                        bar_sync(BAR_FITREQ + sw, NTHREADS);              // A(i) posted
                        do_fit(i % C::NSLOT, fitrows[sw]);
                    }
                    if (has_tail) {
                        bar_sync(BAR_TAILREQ + sw, NTHREADS);             // B(i-2) (or the no-carrier shortcut) posted
                        do_tail(i - 2, i & 1);
                    }
                    if (has_fit) bar_arrive(BAR_FITDONE + sw, NTHREADS);
                }
            }
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::WORKER_REGS));
    }

    // =====================================================================================
    // main (FFT worker) threads
    // =====================================================================================
    // FFT buffer access: shared memory (padded rows) or, for GMEM, L2-only global scratch.
    auto ld8 = [&](uint32_t byte_off) -> float2 {
        if constexpr (GMEM) return __ldcg(reinterpret_cast<const float2 *>(bufc + byte_off));
        else return *reinterpret_cast<const float2 *>(bufc + byte_off);
    };
    auto st8 = [&](uint32_t byte_off, float2 v) {
        if constexpr (GMEM) __stcg(reinterpret_cast<float2 *>(bufc + byte_off), v);
        else *reinterpret_cast<float2 *>(bufc + byte_off) = v;
    };
    // logical element e -> byte offset: GMEM plain (8e); shared: row (e >> 4) of 136 bytes.
    // pass 1: item j, element k1       e = k1*M + j          -> a1_base(j) + k1 * A1_STEP
    // pass 2: item (k1,n3), element n2 e = k1*M + n2*R3 + n3 -> a2_base(k1,n3) + n2 * A2_STEP
    // pass 3: item g, element n3       e = g*R3 + n3         -> a3_base(g) + n3 * 8
    constexpr uint32_t A1_STEP = GMEM ? (uint32_t)M * 8u : (uint32_t)(M / 16) * 136u;
    constexpr uint32_t A2_STEP = GMEM ? (uint32_t)R3 * 8u : 136u;
    auto a1_base = [&](int j) -> uint32_t {
        return GMEM ? (uint32_t)j * 8u : (uint32_t)(j >> 4) * 136u + (uint32_t)(j & 15) * 8u;
    };
    auto a2_base = [&](int k1, int n3) -> uint32_t {
        return GMEM ? (uint32_t)(k1 * M + n3) * 8u : (uint32_t)k1 * A1_STEP + (uint32_t)n3 * 8u;
    };
    auto a3_base = [&](int g) -> uint32_t { return (uint32_t)g * (uint32_t)C::ROW_BYTES; };

    float2 w1[I1], w4[I1];                                      // W_N^j and W_N^{4j}
#pragma unroll
    for (int i = 0; i < I1; ++i) {
        const int j = tid + T * i;
        w1[i] = cispi(-2.0f * (float)j / (float)N);
        w4[i] = cispi(-2.0f * (float)((4 * j) & (N - 1)) / (float)N);
    }

    // thread 0: TMA bulk copy of block i's raw tile into raw stage `stage` (completion on mbar[stage]).
    // FASTDET flow: stage = block parity, two blocks of read-ahead.  Detect flow: stage 0 feeds stage A (FFT#1),
    // stage 1 feeds stage B (mix + FFT#2) FIT_DEPTH blocks later -- the tile is fetched twice (the second time from L2),
    // so each stage is refilled with its next block as soon as pass 1 has consumed it.
    auto issue_tile = [&](int i, int stage) {
        const int blk = (int)blockIdx.x + i * (int)gridDim.x;
        mbar_expect_tx(&mbar[stage], RAW_BYTES);
        tma_bulk_g2s(raw_s + (size_t)stage * RAW_BYTES, p.raw + (size_t)blk * (size_t)p.raw_stride, RAW_BYTES, &mbar[stage]);
    };
    if (use_raw && tid == 0) {
        if (has_block(0)) issue_tile(0, 0);
#ifdef THR_EXP_NOREFETCH
        if (FASTDET && has_block(1)) issue_tile(1, 1);
#else
        if (FASTDET ? has_block(1) : has_block(0)) issue_tile(FASTDET ? 1 : 0, 1);
#endif
    }
    uint32_t par0 = 0, par1 = 0;    // phase parity of the two tile barriers

    // forward passes 1 and 2 of block i (shared by FFT#1 and FFT#2)
    // zoom (pruned FFT#1): only the 128 bins [b0, b0+128) are needed; after the pre-shift by -b0 they are the bins
    // k3 == 0, k2 < 4, so pass 2 computes 4 of its R2 outputs and pass 3 degenerates to a sum; the spectrum energy
    // comes from Parseval
    const bool zoom = !FASTDET && C::ZOOM_OK && (p.zoom != 0) && (p.dbg_fft_mag == nullptr);
    //   shift : multiply the samples by a phasor exp(-2 pi i s n/N) = rho[n1] * ph0[j] (the mix of stage B, or the
    //           integer pre-shift that moves an arbitrary narrow carrier window into the 128-bin zoom band)
    //   stageB: FFT#2 (full pass 2, raw tile released after pass 1); otherwise FFT#1 (pruned if `zoom`)
    //   stage / next: raw stage holding block i, and the block to fetch into it once pass 1 has consumed it
    auto fwd_pass12 = [&](int i, int stage, int next, bool shift, bool stageB, const float2 (&ph0)[I1], const float2 *rho,
                          float &energy) {
        const uint16_t *rawt = reinterpret_cast<const uint16_t *>(raw_s + (size_t)stage * RAW_BYTES);
        const float2 *iqb = use_raw ? nullptr : p.iq + (size_t)((int)blockIdx.x + i * (int)gridDim.x) * N;
        // pass 1: samples (rawconv or complex64) [* mix phasor] -> radix-32 over n1 (stride M)
        //         -> twiddle W_N^{j k1} -> in-place store
#pragma unroll
        for (int it = 0; it < I1; ++it) {
            const int j = tid + T * it;
            float2 x[32];
            if (use_raw) {
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) x[brev(n1, 5)] = rawconv(rawt[n1 * M + j]);
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) x[brev(n1, 5)] = __ldg(&iqb[n1 * M + j]);
            }
            if (!stageB && zoom) {               // sum |x|^2 (Parseval: sum_k |X[k]|^2 = N sum_n |x[n]|^2)
                float2 e2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) e2 = __ffma2_rn(x[n1], x[n1], e2);
                energy += e2.x + e2.y;
            }
            if (shift) {
                // row phasor here; the per-thread phasor ph0 is common to the whole item and is
                // folded into the twiddle seeds below (the DFT is linear)
                const float4 *rho4 = reinterpret_cast<const float4 *>(rho);
#pragma unroll
                for (int n1 = 0; n1 < 32; n1 += 2) {
                    const float4 r = rho4[n1 >> 1];
                    if (n1 > 0) x[brev(n1, 5)] = cmul(x[brev(n1, 5)], make_float2(r.x, r.y));
                    x[brev(n1 + 1, 5)] = cmul(x[brev(n1 + 1, 5)], make_float2(r.z, r.w));
                }
            }
            fft_dit<32, false>(x);
            // Twiddles W_N^{j k1} are regenerated per block from two per-thread seeds.  The opaque
            // asm keeps the compiler from hoisting all 31 powers out of the block loop, which
            // would turn them into a 124 KB per-CTA local-memory table (L2 traffic + latency).
            float2 ws = w1[it], ws4 = w4[it];
            asm volatile("" : "+f"(ws.x), "+f"(ws.y), "+f"(ws4.x), "+f"(ws4.y));
            float2 cur[4];
            cur[0] = ws;
            cur[1] = cmul(ws, ws);
            cur[2] = cmul(cur[1], ws);
            cur[3] = ws4;
            const uint32_t ab = a1_base(j);
            if (shift) {
#pragma unroll
                for (int c = 0; c < 4; ++c) cur[c] = cmul(cur[c], ph0[it]);
                st8(ab, cmul(x[0], ph0[it]));
            } else {
                st8(ab, x[0]);
            }
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) {
                if (k1 > 4) cur[(k1 - 1) & 3] = cmul(cur[(k1 - 1) & 3], ws4);
                st8(ab + (uint32_t)k1 * A1_STEP, cmul(x[k1], cur[(k1 - 1) & 3]));
            }
        }
        bar_sync(BAR_MAIN, T);
        // after the pass-1 barrier nobody reads this raw stage any more: fetch its next block
#ifdef THR_EXP_NOREFETCH
        if (use_raw && tid == 0 && has_block(next) && stage == 0) issue_tile(next, stage);
#else
        if (use_raw && tid == 0 && has_block(next)) issue_tile(next, stage);
#endif
        if constexpr (C::ZOOM_OK) {
            if (!stageB && zoom) {
                // pruned pass 2: outputs k2 = 0..3 of the R2-point DFT over n2 = Q m + r (Q = R2/4):
                //   c_r[k] = sum_m a[Q m + r] W_4^{mk}    (radix-4, no multiplications)
                //   B[k]   = sum_r W_R2^{rk} c_r[k]       (2 packed FMAs per term; W_R2 = W_32^(32/R2))
                constexpr int Q = R2 / 4, E = 32 / R2;
#pragma unroll UNROLL_I2
                for (int it = 0; it < I2; ++it) {
                    const int w = tid + T * it;
                    const int k1 = w >> LOG2R3, n3 = w & (R3 - 1);
                    const uint32_t ab = a2_base(k1, n3);
                    float2 b0, b1, b2, b3;
#pragma unroll
                    for (int r = 0; r < Q; ++r) {
                        const float2 a0 = ld8(ab + (uint32_t)(r) * A2_STEP);
                        const float2 a1 = ld8(ab + (uint32_t)(Q + r) * A2_STEP);
                        const float2 a2 = ld8(ab + (uint32_t)(2 * Q + r) * A2_STEP);
                        const float2 a3 = ld8(ab + (uint32_t)(3 * Q + r) * A2_STEP);
                        const float2 s0 = f2add(a0, a2), s1 = f2sub(a0, a2);
                        const float2 s2 = f2add(a1, a3), s3 = f2sub(a1, a3);
                        const float2 c0 = f2add(s0, s2), c2 = f2sub(s0, s2);
                        const float2 c1 = f2add(s1, rot_mj(s3)), c3 = f2add(s1, rot_pj(s3));
                        if (r == 0) {
                            b0 = c0; b1 = c1; b2 = c2; b3 = c3;
                        } else {
                            b0 = f2add(b0, c0);
                            if (r == 1) { b1 = fma_tw32<1 * E>(b1, c1); b2 = fma_tw32<2 * E>(b2, c2); b3 = fma_tw32<3 * E>(b3, c3); }
                            if (r == 2) { b1 = fma_tw32<2 * E>(b1, c1); b2 = fma_tw32<4 * E>(b2, c2); b3 = fma_tw32<6 * E>(b3, c3); }
                            if (r == 3) { b1 = fma_tw32<3 * E>(b1, c1); b2 = fma_tw32<6 * E>(b2, c2); b3 = fma_tw32<9 * E>(b3, c3); }
                            if (r == 4) { b1 = fma_tw32<4 * E>(b1, c1); b2 = fma_tw32<8 * E>(b2, c2); b3 = fma_tw32<12 * E>(b3, c3); }
                            if (r == 5) { b1 = fma_tw32<5 * E>(b1, c1); b2 = fma_tw32<10 * E>(b2, c2); b3 = fma_tw32<15 * E>(b3, c3); }
                            if (r == 6) { b1 = fma_tw32<6 * E>(b1, c1); b2 = fma_tw32<12 * E>(b2, c2); b3 = fma_tw32<18 * E>(b3, c3); }
                            if (r == 7) { b1 = fma_tw32<7 * E>(b1, c1); b2 = fma_tw32<14 * E>(b2, c2); b3 = fma_tw32<21 * E>(b3, c3); }
                        }
                    }
                    st8(ab, b0);
                    st8(ab + 1u * A2_STEP, cmul(b1, tw2[1 * R3 + n3]));
                    st8(ab + 2u * A2_STEP, cmul(b2, tw2[2 * R3 + n3]));
                    st8(ab + 3u * A2_STEP, cmul(b3, tw2[3 * R3 + n3]));
                }
                if constexpr (C::ZOOM_WARPLOCAL) __syncwarp();   // pruned pass 3 reads this warp's own slabs
                else bar_sync(BAR_MAIN, T);
                return;
            }
        }
        // pass 2: radix-R2 over n2 (stride R3) inside each k1 slab, twiddle W_M^{n3 k2}
        if (R2 > 1) {
#pragma unroll UNROLL_I2
            for (int it = 0; it < I2; ++it) {
                const int w = tid + T * it;
                const int k1 = w >> LOG2R3, n3 = w & (R3 - 1);
                const uint32_t ab = a2_base(k1, n3);
                float2 x[R2];
#pragma unroll
                for (int n2 = 0; n2 < R2; ++n2) x[brev(n2, LOG2R2)] = ld8(ab + (uint32_t)n2 * A2_STEP);
                fft_dit<R2, false>(x);
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) {
                    float2 v = x[k2];
                    if (k2 > 0 && !(TW3 && stageB)) v = cmul(v, tw2[k2 * R3 + n3]);
                    st8(ab + (uint32_t)k2 * A2_STEP, v);
                }
            }
            if constexpr (C::WL23) __syncwarp();           // pass 3 reads this warp's own slabs
            else bar_sync(BAR_MAIN, T);
        }
    };

    // ---- correlation stage for one template (soa_estimator.py:97-143): this thread's pass-3 outputs X'
    // (supplied by get_x) x conj(T)/N -> inverse passes 3', 2', 1' -> |c|^2 windowed arg-max -> TailSlot
    // TW3: powers w^1..w^4 of w = W_M^{k2} for the pass-3 item g (k2 = g mod R2); cur3w4 = w^4 steps the chains
    float2 cur3w4 = make_float2(1.f, 0.f);
    auto tw3_seed = [&](int g, float2 (&cur)[4]) {
        float2 w = cispi(-2.0f * (float)(g & (R2 - 1)) / (float)M);
        asm volatile("" : "+f"(w.x), "+f"(w.y));       // not hoisted out of the block loop (register pressure)
        cur[0] = w;
        cur[1] = cmul(w, w);
        cur[2] = cmul(cur[1], w);
        cur[3] = cmul(cur[1], cur[1]);
        cur3w4 = cur[3];
    };
    int corr_blk = 0;                // block being correlated (offset of the optional per-block correlation output)
    auto corr_stage = [&](int q, int tpl, auto &&get_tv, auto &&get_x) {
#pragma unroll UNROLL_I3
        for (int it = 0; it < I3; ++it) {
            const int g = C::p3_item(tid, it);
            const uint32_t ab = a3_base(g);
            float2 tv[R3];                                    // template spectrum, issued early
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) tv[k3] = get_tv(it, g, k3);
            float2 x[R3];
            get_x(it, g, ab, x);
            // multiply by conj(T)/N (soa_estimator.py:99) and run the inverse radix-R3 DFT
            float2 y[R3];
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) y[brev(k3, LOG2R3)] = cmul(x[k3], tv[k3]);
            fft_dit<R3, true>(y);
            if constexpr (TW3) {      // conj(W_M^{n3 k2}) here instead of on the pass-2' loads
                float2 cur[4];
                tw3_seed(g, cur);
#pragma unroll
                for (int n3 = 1; n3 < R3; ++n3) {
                    if (n3 > 4) cur[(n3 - 1) & 3] = cmul(cur[(n3 - 1) & 3], cur3w4);
                    y[n3] = cmulc(y[n3], cur[(n3 - 1) & 3]);
                }
            }
#pragma unroll
            for (int n3 = 0; n3 < R3; ++n3) st8(ab + (uint32_t)n3 * 8u, y[n3]);
        }
        if constexpr (C::WL23 && R2 > 1) __syncwarp();   // pass 2' reads this warp's own slabs
        else bar_sync(BAR_MAIN, T);
        // inverse pass 2': conj twiddle on load, radix-R2 over k2
        if (R2 > 1) {
#pragma unroll UNROLL_I2
            for (int it = 0; it < I2; ++it) {
                const int w = tid + T * it;
                const int k1 = w >> LOG2R3, n3 = w & (R3 - 1);
                const uint32_t ab = a2_base(k1, n3);
                float2 x[R2];
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) {
                    float2 v = ld8(ab + (uint32_t)k2 * A2_STEP);
                    if (k2 > 0 && !TW3) v = cmulc(v, tw2[k2 * R3 + n3]);
                    x[brev(k2, LOG2R2)] = v;
                }
                fft_dit<R2, true>(x);
#pragma unroll
                for (int n2 = 0; n2 < R2; ++n2) st8(ab + (uint32_t)n2 * A2_STEP, x[n2]);
            }
            bar_sync(BAR_MAIN, T);
        }
        // inverse pass 1': conj twiddle on load, radix-32 over k1 -> c[n1*M + j]
        float cp[I1][32];
        float c1sum = 0.f, c2sum = 0.f;
        float cbestv = 0.f;                  // best in-window |c|^2 of this thread
        uint32_t inmask[I1];                 // bit n1: lag n1*M + j lies in [corr_start, corr_stop)
#pragma unroll
        for (int it = 0; it < I1; ++it) {
            const int j = tid + T * it;
            const uint32_t ab = a1_base(j);
            float2 x[32];
            float2 ws = w1[it], ws4 = w4[it];
            asm volatile("" : "+f"(ws.x), "+f"(ws.y), "+f"(ws4.x), "+f"(ws4.y));
            float2 cur[4];
            cur[0] = ws;
            cur[1] = cmul(ws, ws);
            cur[2] = cmul(cur[1], ws);
            cur[3] = ws4;
            x[0] = ld8(ab);
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) {
                if (k1 > 4) cur[(k1 - 1) & 3] = cmul(cur[(k1 - 1) & 3], ws4);
                x[brev(k1, 5)] = cmulc(ld8(ab + (uint32_t)k1 * A1_STEP), cur[(k1 - 1) & 3]);
            }
            fft_dit<32, true>(x);
            // |c|^2 and windowed arg-max over [corr_start, corr_stop) (soa_estimator.py:137-143);
            // n = n1*M + j grows with n1, so '>' keeps the first maximum
            {
                // rows n1 in [lo, hi] are inside the window for this thread's column j
                const int lo = max(0, (p.corr_start - j + M - 1) >> LOG2M);
                const int hi = min(31, (p.corr_stop - 1 - j) >> LOG2M);    // -1 if j >= corr_stop
                const uint32_t upto_hi = hi >= 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u);
                inmask[it] = (hi >= lo) ? (upto_hi & ~((1u << lo) - 1u)) : 0u;
            }
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const float pv = x[n1].x * x[n1].x + x[n1].y * x[n1].y;
                cp[it][n1] = pv;
                if (inmask[it] & (1u << n1)) cbestv = fmaxf(cbestv, pv);
            }
            if (need_std_k) {
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    if (n1 * M + j < p.corr_len) {
                        c1sum += sqrtf(cp[it][n1]);
                        c2sum += cp[it][n1];
                    }
                }
            }
            if (p.dbg_corr && tpl == 0) {
                float2 *corr_out = p.dbg_corr + (size_t)corr_blk * (size_t)p.dbg_corr_stride;
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1)
                    if (n1 * M + j < p.corr_len) corr_out[n1 * M + j] = x[n1];
            }
        }
        // block arg-max: first maximum of |c|^2 over the window (soa_estimator.py:137-143)
        auto find_lag = [&](uint32_t gb) {
            uint32_t key = 0xffffffffu;
#pragma unroll
            for (int it = 0; it < I1; ++it) {
#pragma unroll
                for (int n1 = 31; n1 >= 0; --n1)
                    if ((inmask[it] & (1u << n1)) && __float_as_uint(cp[it][n1]) == gb)
                        key = min(key, (uint32_t)(n1 * M + tid + T * it));
            }
            return key;
        };
        TailSlot &ts = tailslot[q * TPL_SLOTS + tpl];
        ArgOut rb;
        if (need_std_k) rb = main_argmax<T, true>(__float_as_uint(cbestv), c1sum, c2sum, red, tid, find_lag);
        else rb = main_argmax<T, false>(__float_as_uint(cbestv), 0.f, 0.f, red, tid, find_lag);
        const int s = (int)rb.key;
        if (tid == 0) {
            ts.peak_cp = __uint_as_float(rb.vbits);
            ts.s = s;
            ts.c1 = rb.s0;
            ts.c2 = rb.s1;
        }
        // neighbours of the peak for the Gaussian interpolation
#pragma unroll
        for (int it = 0; it < I1; ++it) {
            const int j = tid + T * it;
#pragma unroll
            for (int dd = -1; dd <= 1; dd += 2) {
                const int nt = s + dd;
                if (nt >= 0 && nt < N && (nt & (M - 1)) == j) {
                    const int n1s = nt >> LOG2M;
                    float v = 0.f;
#pragma unroll
                    for (int n1 = 0; n1 < 32; ++n1) v = (n1 == n1s) ? cp[it][n1] : v;
                    if (dd < 0) ts.pa = v; else ts.pc = v;
                }
            }
        }
    };

    // =====================================================================================
    // FASTDET: the reference's native semantics (fastcard/fastcard.c:177-189, cardet.c:7-41,
    // fastdet/corr_detector.cpp:127-197).  Per block: FFT#1 (full) -> |X|^2, sum, windowed arg-max,
    // power-domain threshold -> X x conj(T)[k - kpeak]/N -> inverse transform -> |c|^2 arg-max -> tail.
    // The integer roll of the spectrum is a re-indexing of the template, so X never leaves the registers
    // between the forward pass 3 and the inverse pass 3'.
    // =====================================================================================
    if constexpr (FASTDET) {
        for (int i = 0; has_block(i); ++i) {
            const int q = i & 1;
            if (use_raw) {
                mbar_wait(&mbar[q], q ? par1 : par0);
                if (q) par1 ^= 1; else par0 ^= 1;
            }
            float2 ph_unused[I1];
#pragma unroll
            for (int it = 0; it < I1; ++it) ph_unused[it] = make_float2(1.f, 0.f);
            float tenergy = 0.f;
            fwd_pass12(i, q, i + 2, false, false, ph_unused, nullptr, tenergy);
            // pass 3, power spectrum (fastcard.c:180), sum (cardet.c:12) and windowed maximum (cardet.c:15-19)
            float2 xk[I3][R3];
            float esum = 0.f, bestv = 0.f;
            const bool all_in = (p.win_len >= N);
#pragma unroll
            for (int it = 0; it < I3; ++it) {
                const int g = C::p3_item(tid, it);
                const uint32_t ab = a3_base(g);
#pragma unroll
                for (int n3 = 0; n3 < R3; ++n3) xk[it][brev(n3, LOG2R3)] = ld8(ab + (uint32_t)n3 * 8u);
                fft_dit<R3, false>(xk[it]);
                const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
                const uint32_t relb = (uint32_t)(kb - p.win_start) & (uint32_t)(N - 1);
                const bool item_in = all_in || ((relb & (uint32_t)(S - 1)) < (uint32_t)p.win_len);
#pragma unroll
                for (int k3 = 0; k3 < R3; ++k3) {
                    const float pv = xk[it][k3].x * xk[it][k3].x + xk[it][k3].y * xk[it][k3].y;
                    esum += pv;
                    if (item_in) {
                        const uint32_t rel = (relb + (uint32_t)(S * k3)) & (uint32_t)(N - 1);
                        if (rel < (uint32_t)p.win_len) bestv = fmaxf(bestv, pv);
                    }
                    if (p.dbg_fft_mag) p.dbg_fft_mag[kb + S * k3] = sqrtf(pv);
                }
            }
            const ArgOut ra = main_argmax<T, true>(__float_as_uint(bestv), esum, 0.f, red, tid, [&](uint32_t gb) {
                uint32_t key = 0xffffffffu;
#pragma unroll
                for (int it = 0; it < I3; ++it) {
                    const int g = C::p3_item(tid, it);
                    const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
                    const uint32_t relb = (uint32_t)(kb - p.win_start) & (uint32_t)(N - 1);
#pragma unroll
                    for (int k3 = 0; k3 < R3; ++k3) {
                        const float pv = xk[it][k3].x * xk[it][k3].x + xk[it][k3].y * xk[it][k3].y;
                        const uint32_t rel = (relb + (uint32_t)(S * k3)) & (uint32_t)(N - 1);
                        if (rel < (uint32_t)p.win_len && __float_as_uint(pv) == gb) key = min(key, rel);
                    }
                }
                return key;
            });
            // cardet.c:21-29: noise power, power-domain threshold
            const float peak_pw = __uint_as_float(ra.vbits), fsum = ra.s0;
            const int kpeak = (p.win_start + (int)ra.key) & (N - 1);
            const float noise_pw = (fsum != 0.f) ? (fsum - 2.f * peak_pw) / (float)(N - 1) : 0.f;
            const bool carrier = peak_pw > p.c_const + p.c_snr * noise_pw;
            // mailbox q was last used by block i-2: wait until its tail has been written out
            if constexpr (SERVICE) {
                if (i >= 2) bar_sync(BAR_FITDONE + q, NTHREADS);
            }
            TailHdr &h = tailhdr[q];
            if (tid == 0) {
                h.kpeak = kpeak;
                h.carrier = carrier ? 1 : 0;
                h.peak_mag = peak_pw;
                h.noise_c = noise_pw;
                h.sig_energy1 = fsum / (float)N;          // corr_detector.cpp:184
                h.delta = 0.f;
            }
            if (carrier) {
                // |X|^2 next to the peak for the parabolic carrier offset (corr_detector.cpp:191)
#pragma unroll
                for (int it = 0; it < I3; ++it) {
                    const int g = C::p3_item(tid, it);
                    const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
                    const uint32_t u = (uint32_t)(kb - kpeak + 1) & (uint32_t)(N - 1);
                    const uint32_t lo = u & (uint32_t)(S - 1);
                    if (lo < 3u && lo != 1u) {
                        const int k3s = (R3 - (int)(u >> LOG2S)) & (R3 - 1);
                        float v = 0.f;
#pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3)
                            v = (k3 == k3s) ? xk[it][k3].x * xk[it][k3].x + xk[it][k3].y * xk[it][k3].y : v;
                        h.pad[lo >> 1] = v;
                    }
                }
                const int rel = (int)ra.key;
                auto keep_x = [&](int it, int, uint32_t, float2 (&x)[R3]) {
#pragma unroll
                    for (int k3 = 0; k3 < R3; ++k3) x[k3] = xk[it][k3];
                };
                if (p.tpl_shift) {      // pre-rolled template spectrum of this carrier bin: coalesced, base + immediate
                    const float2 *tsp = p.tpl_shift + (size_t)rel * N + tid;
                    corr_stage(q, 0, [&](int it, int, int k3) { return __ldg(&tsp[(it * R3 + k3) * T]); }, keep_x);
                } else {                // window too wide for a table: gather from the natural-order spectrum
                    corr_stage(q, 0,
                               [&](int, int g, int k3) {
                                   const int k = (g >> LOG2R2) + 32 * (g & (R2 - 1)) + S * k3;
                                   return __ldg(&p.tpl_nat[(k - kpeak) & (N - 1)]);
                               },
                               keep_x);
                }
            }
            if constexpr (SERVICE) {
                bar_arrive(BAR_TAILREQ + q, NTHREADS);
            } else {
                bar_sync(BAR_MAIN, T);
                if (tid < 32) do_tail_fd(i, q);
            }
        }
        return;
    }

    constexpr int FIT_DEPTH = C::FIT_DEPTH, NSLOT = C::NSLOT;
    if constexpr (!FASTDET) for (int i = -FIT_DEPTH; has_block(i < 0 ? 0 : i); ++i) {
        // ================================================================= A(i+FIT_DEPTH): FFT #1
        if (has_block(i + FIT_DEPTH)) {
            const int ia = i + FIT_DEPTH, q = ia % NSLOT;
            if constexpr (STAGES == 2) {
                // no stage A: the caller supplies the shifted spectrum.  The mailbox protocol stays (slot says "go")
                if (tid == 0) {
                    FitSlot &fs0 = fitslot[q];
                    fs0.kpeak = 0;
                    fs0.carrier = 1;
                    fs0.peak_mag = 0.f;
                    fs0.noise_c = 0.f;
                    fs0.sig_energy1 = 0.f;
                    fs0.delta = 0.f;
                }
            } else {
            if (use_raw) {
                mbar_wait(&mbar[0], par0);
                par0 ^= 1;
            }
            // zoom band not at bin 0: pre-shift the block by -b0 bins, folded into pass 1 like the mix of stage B
            const bool shiftA = zoom && (p.zoom_base != 0);
            float2 phA[I1];
#pragma unroll
            for (int it = 0; it < I1; ++it) {
                phA[it] = make_float2(1.f, 0.f);
                if (shiftA) {
                    const int e = (int)(((long long)p.zoom_base * (tid + T * it)) & (N - 1));
                    phA[it] = cispi(-2.0f * (float)e / (float)N);
                }
            }
            float tenergy = 0.f;
            if constexpr (ROLL) {
                fwd_pass12(ia, 0, ia + 1, shiftA, false, phA, zrho, tenergy);               // one copy: code size first
            } else {
                if (shiftA) fwd_pass12(ia, 0, ia + 1, true, false, phA, zrho, tenergy);     // two specialised copies: a run-time
                else fwd_pass12(ia, 0, ia + 1, false, false, phA, zrho, tenergy);           // flag inside pass 1 costs registers
            }

            FitSlot &fs = fitslot[q];                   // ring slot of block ia
            // carrier decision in float32 (carrier_detect.py:99-115); returns the peak bin or -1
            auto decide = [&](const ArgOut &ra) -> int {
                const float peak_pw = __uint_as_float(ra.vbits);
                const int kpeak = (p.win_start + (int)ra.key) & (N - 1);
                const float peak_mag = sqrtf(peak_pw);
                const float noise_pw_c = (ra.s0 - 2.f * (peak_mag * peak_mag)) / (float)(N - 1);
                const float noise_c = sqrtf(noise_pw_c);
                float var_c = 0.f;
                if (need_std_c) {
                    const float mean = ra.s1 / (float)N;
                    var_c = fmaxf(ra.s0 / (float)N - mean * mean, 0.f);
                }
                const float thr_c = sqrtf(p.c_const + p.c_snr * (noise_c * noise_c) + p.c_std * var_c);
                const bool carrier = peak_mag > thr_c;
                if (tid == 0) {
                    fs.kpeak = kpeak;
                    fs.carrier = carrier ? 1 : 0;
                    fs.peak_mag = peak_mag;
                    fs.noise_c = noise_c;
                    fs.sig_energy1 = ra.s0 / (float)N;
                    fs.delta = 0.f;
                }
                return carrier ? kpeak : -1;
            };
            if (zoom) {
                ArgOut ra;
                // pass 3 (pruned): X[k1 + 32 k2] = sum_n3 B[k1,k2;n3] for the 128 bins k < 128;
                // 4 threads per bin, R3/4 terms each, then two shuffles
                constexpr int PER = R3 / 4;
                if constexpr (C::ZOOM_WARPLOCAL) {
                    // warp w owns slabs k1 = 2w, 2w+1 (written by its own pruned pass 2): 8 bins x 4 lanes
                    const int k1 = 2 * (tid >> 5) + (lane >> 4), k2 = (lane >> 2) & 3, sub = lane & 3;
                    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int t = 0; t < PER; ++t)
                        acc = f2add(acc, ld8(a2_base(k1, sub * PER + t) + (uint32_t)k2 * A2_STEP));
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
                    if (sub == 0) zpow[k1 + 32 * k2] = acc.x * acc.x + acc.y * acc.y;
                } else {
#pragma unroll 1
                    for (int pzb = 0; pzb < 128; pzb += T / 4) {
                        const int pz = pzb + (tid >> 2), sub = tid & 3;
                        const int k1 = pz & 31, k2 = pz >> 5;
                        float2 acc = make_float2(0.f, 0.f);
                        if (pz < 128) {
#pragma unroll
                            for (int t = 0; t < PER; ++t)
                                acc = f2add(acc, ld8(a2_base(k1, sub * PER + t) + (uint32_t)k2 * A2_STEP));
                        }
                        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
                        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
                        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
                        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
                        if (sub == 0 && pz < 128) zpow[k1 + 32 * k2] = acc.x * acc.x + acc.y * acc.y;
                    }
                }
                const float wsum = warp_sum(tenergy);
                if (lane == 0) red[16 + (tid >> 5)] = __float_as_uint(wsum);
                bar_sync(BAR_MAIN, T);
                // every warp finds the window maximum of the 128 powers on its own (no extra barrier)
                uint32_t vb = 0u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k = lane + 32 * c;
                    const uint32_t rel = (uint32_t)(k - p.zoom_w0);
                    vb = max(vb, rel < (uint32_t)p.win_len ? __float_as_uint(zpow[k]) : 0u);
                }
                const uint32_t gb = __reduce_max_sync(0xffffffffu, vb);
                uint32_t key = 0xffffffffu;
#pragma unroll
                for (int c = 3; c >= 0; --c) {
                    const int k = lane + 32 * c;
                    const uint32_t rel = (uint32_t)(k - p.zoom_w0);
                    if (rel < (uint32_t)p.win_len && __float_as_uint(zpow[k]) == gb) key = rel;
                }
                ra.vbits = gb;
                ra.key = __reduce_min_sync(0xffffffffu, key);
                float esum = 0.f;
#pragma unroll
                for (int w = 0; w < T / 32; ++w) esum += __uint_as_float(red[16 + w]);
                ra.s0 = esum * (float)N;
                ra.s1 = 0.f;
                const int kpeak = decide(ra);
                // the 7 magnitudes around the peak for the Dirichlet fit
                if (kpeak >= 0 && tid < 7) fs.mags[tid] = sqrtf(zpow[p.zoom_w0 + (int)ra.key - 3 + tid]);
            } else {
                // pass 3 + power spectrum (Signal.mag, signal_utils.py:99-107) + windowed arg-max
                float pw[I3][R3];
                float esum = 0.f, msum = 0.f;
                float bestv = 0.f;                       // best in-window power of this thread
                const bool all_in = (p.win_len >= N);    // default window '0--1': every bin qualifies
    #pragma unroll
                for (int it = 0; it < I3; ++it) {
                    const int g = C::p3_item(tid, it);
                    const uint32_t ab = a3_base(g);
                    float2 x[R3];
    #pragma unroll
                    for (int n3 = 0; n3 < R3; ++n3) x[brev(n3, LOG2R3)] = ld8(ab + (uint32_t)n3 * 8u);
                    fft_dit<R3, false>(x);
                    const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));     // k1 + 32 k2
    #pragma unroll
                    for (int k3 = 0; k3 < R3; ++k3) {
                        const int r = k3;
                        const float pv = x[r].x * x[r].x + x[r].y * x[r].y;
                        pw[it][k3] = pv;
                        esum += pv;
                    }
                    if (need_std_c) {
    #pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3) msum += sqrtf(pw[it][k3]);
                    }
                    // window test: bin k = kb + S*k3, rel = (k - win_start) mod N must be < win_len.
                    // rel mod S does not depend on k3, so a narrow window rejects most items at once.
                    const uint32_t relb = (uint32_t)(kb - p.win_start) & (uint32_t)(N - 1);
                    if (all_in) {
    #pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3) bestv = fmaxf(bestv, pw[it][k3]);
                    } else if ((relb & (uint32_t)(S - 1)) < (uint32_t)p.win_len) {
    #pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3) {
                            const uint32_t rel = (relb + (uint32_t)(S * k3)) & (uint32_t)(N - 1);
                            if (rel < (uint32_t)p.win_len) bestv = fmaxf(bestv, pw[it][k3]);
                        }
                    }
                    if (p.dbg_fft_mag) {
    #pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3) p.dbg_fft_mag[kb + S * k3] = sqrtf(pw[it][k3]);
                    }
                }
                // first maximum in window order (np.argmax over the wrapped window, carrier_detect.py:138-149)
                const ArgOut ra = main_argmax<T, true>(__float_as_uint(bestv), esum, msum, red, tid, [&](uint32_t gb) {
                    uint32_t key = 0xffffffffu;
    #pragma unroll
                    for (int it = 0; it < I3; ++it) {
                        const int g = C::p3_item(tid, it);
                        const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
                        const uint32_t relb = (uint32_t)(kb - p.win_start) & (uint32_t)(N - 1);
    #pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3) {
                            const uint32_t rel = (relb + (uint32_t)(S * k3)) & (uint32_t)(N - 1);
                            if (rel < (uint32_t)p.win_len && __float_as_uint(pw[it][k3]) == gb) key = min(key, rel);
                        }
                    }
                    return key;
                });
                const int kpeak = decide(ra);
                // the 7 magnitudes around the peak for the Dirichlet fit (compare-select, no dynamic
                // register indexing)
                if (kpeak >= 0) {
#pragma unroll
                    for (int it = 0; it < I3; ++it) {
                        const int g = C::p3_item(tid, it);
                        const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
                        const uint32_t u = (uint32_t)(kb - kpeak + 3) & (uint32_t)(N - 1);
                        const uint32_t lo = u & (uint32_t)(S - 1);
                        if (lo < 7u) {
                            const int k3s = (R3 - (int)(u >> LOG2S)) & (R3 - 1);
                            float v = 0.f;
#pragma unroll
                            for (int k3 = 0; k3 < R3; ++k3) v = (k3 == k3s) ? pw[it][k3] : v;
                            fs.mags[lo] = sqrtf(v);
                        }
                    }
                }
            }
            }   // STAGES != 2
            if constexpr (SERVICE) {
                // that fit warp: fit(ia) may start.  Once the pipeline is full the request is posted a few instructions
                // later, behind the FITDONE wait of stage B below: with FIT_DEPTH == NFITW it goes to the warp whose
                // previous fit that wait is for, so the warp can never be offered two requests at once.
                if (i < 0) bar_arrive(BAR_FITREQ + ia % C::NFITW, NTHREADS);
            } else {
                bar_sync(BAR_MAIN, T);
                if (tid < 32) do_fit(q, fitrows[0]);   // visible to all after the next BAR_MAIN
            }
        }

        // ================================================================= B(i): mix, FFT #2, correlation
        if (i >= 0) {
            const int q = i & 1;                        // tail mailbox
            if constexpr (SERVICE) {
                bar_sync(BAR_FITDONE + i % C::NFITW, NTHREADS);           // fit(i) finished, tail(i-2) has read mailbox q
                if (has_block(i + FIT_DEPTH)) bar_arrive(BAR_FITREQ + (i + FIT_DEPTH) % C::NFITW, NTHREADS);
            } else {
                bar_sync(BAR_MAIN, T);
            }
#ifndef THR_EXP_NOREFETCH                               // (timing experiment: stage B on whatever stage 1 holds)
            if (use_raw) {                              // the re-fetched raw tile of block i (stage 1)
                mbar_wait(&mbar[1], par1);
                par1 ^= 1;
            }
#endif
            const FitSlot &fs = fitslot[i % NSLOT];
            const int kpeak = fs.kpeak;
            const bool carrier = fs.carrier != 0;
            if (tid == 0) {
                TailHdr &h = tailhdr[q];
                h.kpeak = kpeak;
                h.carrier = fs.carrier;
                h.peak_mag = fs.peak_mag;
                h.noise_c = fs.noise_c;
                h.sig_energy1 = fs.sig_energy1;
                h.delta = fs.delta;
            }
            if (!carrier) {
                // the mix does not run: stage 1 goes straight to the next block -- once every thread has seen this
                // block's tile arrive (a thread that is still to wait on the tile barrier must not find it a phase ahead)
                if (use_raw) {
                    bar_sync(BAR_MAIN, T);
                    if (tid == 0 && has_block(i + 1)) issue_tile(i + 1, 1);
                }
                if constexpr (SERVICE) {
                    bar_arrive(BAR_TAILREQ + (i + 2) % C::NFITW, NTHREADS);   // the warp that will fit block i + 2
                } else {
                    bar_sync(BAR_MAIN, T);
                    if (tid < 32) do_tail(i, q);
                }
                continue;
            }
            const float delta = fs.delta;

            const int blk_b = (int)blockIdx.x + i * (int)gridDim.x;
            corr_blk = blk_b;
            float2 *sfft_out = p.dbg_shifted_fft ? p.dbg_shifted_fft + (size_t)blk_b * (size_t)p.dbg_sfft_stride : nullptr;
            // ---- mix + FFT #2 (carrier_sync.py:222-238)
            // x'[n] = x[n] exp(2 pi i shift (n/N - 1/2)), shift = -(k + delta); n = n1*M + j
            if constexpr (STAGES != 2) {
                float2 ph0[I1];
#pragma unroll
                for (int it = 0; it < I1; ++it) {
                    const int j = tid + T * it;
                    const int e = (int)(((long long)kpeak * j) & (N - 1));
                    float turns = -((float)e / (float)N) - delta * ((float)j / (float)N);
                    turns += 0.5f * (float)(kpeak & 1) + 0.5f * delta;
                    ph0[it] = cispi(2.f * turns);
                }
                float unused_energy = 0.f;
                fwd_pass12(i, 1, i + 1, true, true, ph0, fs.rho, unused_energy);
            }
            // pass 3 of FFT#2 for one item: this thread's R3 outputs X'[kb + S k3]
            auto load_x = [&](int it, int g, uint32_t ab, float2 (&x)[R3]) {
                if constexpr (TW3) {    // W_M^{n3 k2} on the loads (pass 2 stored its outputs untwiddled)
                    float2 cur[4];
                    tw3_seed(g, cur);
                    x[0] = ld8(ab);
#pragma unroll
                    for (int n3 = 1; n3 < R3; ++n3) {
                        if (n3 > 4) cur[(n3 - 1) & 3] = cmul(cur[(n3 - 1) & 3], cur3w4);
                        x[brev(n3, LOG2R3)] = cmul(ld8(ab + (uint32_t)n3 * 8u), cur[(n3 - 1) & 3]);
                    }
                } else {
#pragma unroll
                    for (int n3 = 0; n3 < R3; ++n3) x[brev(n3, LOG2R3)] = ld8(ab + (uint32_t)n3 * 8u);
                }
                fft_dit<R3, false>(x);
                if (sfft_out) {
                    const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
#pragma unroll
                    for (int k3 = 0; k3 < R3; ++k3) sfft_out[kb + S * k3] = x[k3];
                }
            };

            if constexpr (STAGES == 1) {
                // the Synchronizer's output is the shifted spectrum: finish FFT#2 and stop
#pragma unroll
                for (int it = 0; it < I3; ++it) {
                    const int g = C::p3_item(tid, it);
                    float2 x[R3];
                    load_x(it, g, a3_base(g), x);
                }
            } else if constexpr (STAGES == 2) {
                // the SoaEstimator's input is a shifted spectrum: X' comes from global memory, and with it the signal
                // energy mean |X'|^2 of the noise estimate (soa_estimator.py:108-120)
                const float2 *src = p.in_sfft + (size_t)blk_b * N;
                float esum = 0.f;
                for (int tpl = 0; tpl < n_tpl; ++tpl) {
                    const float2 *tsp = p.tpl_spec + (size_t)tpl * N;
                    corr_stage(q, tpl, [&](int it, int, int k3) { return ldg_stream(&tsp[(size_t)(it * R3 + k3) * T + tid]); },
                               [&](int, int g, uint32_t, float2 (&x)[R3]) {
                        const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
#pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3) {
                            x[k3] = __ldg(&src[kb + S * k3]);
                            if (tpl == 0) esum += x[k3].x * x[k3].x + x[k3].y * x[k3].y;
                        }
                    });
                }
                esum = warp_sum(esum);
                if (lane == 0) red[16 + (tid >> 5)] = __float_as_uint(esum);
                bar_sync(BAR_MAIN, T);
                if (tid == 0) {
                    float tot = 0.f;
                    for (int w = 0; w < T / 32; ++w) tot += __uint_as_float(red[16 + w]);
                    tailhdr[q].sig_energy1 = tot / (float)N;
                }
            } else {
            // ---- per template: x conj(T)/N and the inverse transform
            for (int tpl = 0; tpl < n_tpl; ++tpl) {
                const float2 *tsp = p.tpl_spec + (size_t)tpl * N;
                corr_stage(q, tpl, [&](int it, int, int k3) { return ldg_stream(&tsp[(size_t)(it * R3 + k3) * T + tid]); },
                           [&](int it, int g, uint32_t ab, float2 (&x)[R3]) {
                if (tpl == 0) {
                    load_x(it, g, ab, x);
                    if (MULTI && p.n_templates > 1) {
#pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3)
                            p.xsave[(size_t)blockIdx.x * N + (size_t)(it * R3 + k3) * T + tid] = x[k3];
                    }
                } else {
#pragma unroll
                    for (int k3 = 0; k3 < R3; ++k3)
#ifdef THR_EXP_NOXLOAD          // timing experiment only (wrong results): what re-reading X' from L2 costs
                        x[k3] = make_float2((float)tid, (float)(k3 + tpl));
#else
                        x[k3] = p.xsave[(size_t)blockIdx.x * N + (size_t)(it * R3 + k3) * T + tid];
#endif
                }
                });
                // next template reuses the FFT buffer: all pass-1' loads are done (reduction barriers)
            }
            }
            if constexpr (SERVICE) {
                bar_arrive(BAR_TAILREQ + (i + 2) % C::NFITW, NTHREADS);   // that service warp: tail(i) may start
            } else {
                bar_sync(BAR_MAIN, T);
                if (tid < 32) do_tail(i, q);
            }
        }
    }
#undef use_raw
#undef need_std_c
#undef need_std_k
}

}  // namespace thr
