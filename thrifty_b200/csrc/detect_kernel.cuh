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
// (shared memory for N <= 16384, XOR-swizzled so that every pass is bank-conflict free;
// an L2-resident global scratch for N = 32768).  Each pass is a radix-32/R2/R3 DFT held
// entirely in registers.  The forward transform leaves the spectrum digit-reversed across
// threads; the template spectrum is stored pre-permuted to match and the inverse transform
// runs the mirrored passes, so no reordering pass exists.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/thrifty_b200.h"

namespace thr {

struct DetectParams {
    const uint8_t *raw;        // [n_blocks][2N] u8 interleaved I,Q (or nullptr)
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
    float fit_tab[7][4];       // per fit point x=-3..3: sin(aWx), cos(aWx), sin(ax), cos(ax), a = pi/N
    float fit_W;               // carrier_len
    float fit_WoverN;          // W / N
    float fit_invN;            // 1 / N
    float2 *dbg_shifted_fft;   // optional [N], natural order (single-block debug launches)
    float2 *dbg_corr;          // optional [corr_len]
    float  *dbg_fft_mag;       // optional [N]
};

template <int LOG2N_, int T_, bool GMEM_>
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
    // position of logical element e in the FFT buffer
    __device__ __forceinline__ static int pos(int e) {
        return GMEM_ ? e : (e ^ ((e >> 4) & 15));
    }
    static constexpr size_t smem_bytes() {
        return (GMEM_ ? 0 : (size_t)N * 8) + 2 * (size_t)(2 * N) + (size_t)M * 8 + 32 * 8 + 1024 + 64;
    }
};

// ------------------------------------------------------------------ small helpers
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// cis(pi * x): cos(pi x) + i sin(pi x), exact range reduction
__device__ __forceinline__ float2 cispi(float x) {
    float s, c;
    sincospif(x, &s, &c);
    return make_float2(c, s);
}
__host__ __device__ constexpr int brev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

// cos / sin of 2*pi*q/32, q in [0,16)
__host__ __device__ constexpr float cos32(int q) {
    return q == 0 ? 1.0f : q == 1 ? 0.98078528040323043f : q == 2 ? 0.92387953251128674f
         : q == 3 ? 0.83146961230254524f : q == 4 ? 0.70710678118654752f
         : q == 5 ? 0.55557023301960218f : q == 6 ? 0.38268343236508977f
         : q == 7 ? 0.19509032201612825f : q == 8 ? 0.0f : -cos32(16 - q);
}
__host__ __device__ constexpr float sin32(int q) { return q <= 8 ? cos32(8 - q) : cos32(q - 8); }

// In-register radix-2 DIF FFT of size R (power of two <= 32), forward sign (e^{-i...}).
// Output k is left at index brev(k).  Calling it as fft_dif(xi, xr) computes the inverse
// (unnormalised) transform, by the swap identity idft(x) = swap(dft(swap(x))).
template <int R>
__device__ __forceinline__ void fft_dif(float (&xr)[R], float (&xi)[R]) {
#pragma unroll
    for (int len = R; len >= 2; len >>= 1) {
        const int half = len >> 1;
#pragma unroll
        for (int base = 0; base < R; base += len) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const int a = base + i, b = a + half;
                const int q = i * (32 / len);            // twiddle W_len^i = W_32^q
                const float ur = xr[a] + xr[b], ui = xi[a] + xi[b];
                const float vr = xr[a] - xr[b], vi = xi[a] - xi[b];
                xr[a] = ur;
                xi[a] = ui;
                if (q == 0) {
                    xr[b] = vr;
                    xi[b] = vi;
                } else if (q == 8) {                     // * (-i)
                    xr[b] = vi;
                    xi[b] = -vr;
                } else if (q == 4) {                     // * (1 - i)/sqrt2
                    const float h = 0.70710678118654752f;
                    xr[b] = (vr + vi) * h;
                    xi[b] = (vi - vr) * h;
                } else if (q == 12) {                    // * (-1 - i)/sqrt2
                    const float h = 0.70710678118654752f;
                    xr[b] = (vi - vr) * h;
                    xi[b] = -(vr + vi) * h;
                } else {
                    const float c = cos32(q), s = sin32(q);   // W = c - i s
                    xr[b] = vr * c + vi * s;
                    xi[b] = vi * c - vr * s;
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
// Least-squares fit of A*|D(x - d)| to 7 magnitudes at x = -3..3 (carrier_sync.py:150-196,
// scipy curve_fit 'lm' from p0 = (y[0], 0)).  Executed by one warp: lane i < 7 owns point i.
// D(z) = sin(aWz) / (W sin(az)), a = pi/N.  The trig of the fixed abscissae comes from
// fit_tab; each iteration only needs sincos of a*W*d and a*d (angle-difference identities).
struct FitSums {
    float jaa, jad, jdd, jar, jdr, cost;
};
__device__ __forceinline__ FitSums fit_eval(float y, bool active, float A, float d, const float (&tab)[4],
                                            float W, float WoverN, float invN) {
    float sd1, cd1, sd2, cd2;
    sincospif(WoverN * d, &sd1, &cd1);
    sincospif(invN * d, &sd2, &cd2);
    const float s1 = tab[0] * cd1 - tab[1] * sd1;     // sin(aW(x-d))
    const float c1 = tab[1] * cd1 + tab[0] * sd1;     // cos(aW(x-d))
    const float s2 = tab[2] * cd2 - tab[3] * sd2;     // sin(a(x-d))
    const float c2 = tab[3] * cd2 + tab[2] * sd2;     // cos(a(x-d))
    float D, Dp;
    if (fabsf(s2) < 1e-30f) {
        D = 1.0f;
        Dp = 0.0f;
    } else {
        const float inv = 1.0f / (W * s2);
        D = s1 * inv;
        const float a = 3.14159265358979f * invN;
        Dp = a * (W * c1 * s2 - s1 * c2) * inv / s2;
    }
    const float g = fabsf(D);
    const float gp = D < 0.f ? -Dp : Dp;
    const float r = y - A * g;
    const float ja = g;
    const float jd = -A * gp;
    FitSums f;
    f.jaa = active ? ja * ja : 0.f;
    f.jad = active ? ja * jd : 0.f;
    f.jdd = active ? jd * jd : 0.f;
    f.jar = active ? ja * r : 0.f;
    f.jdr = active ? jd * r : 0.f;
    f.cost = active ? r * r : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // full warp: keeps every lane's control flow uniform
        f.jaa += __shfl_xor_sync(0xffffffffu, f.jaa, o);
        f.jad += __shfl_xor_sync(0xffffffffu, f.jad, o);
        f.jdd += __shfl_xor_sync(0xffffffffu, f.jdd, o);
        f.jar += __shfl_xor_sync(0xffffffffu, f.jar, o);
        f.jdr += __shfl_xor_sync(0xffffffffu, f.jdr, o);
        f.cost += __shfl_xor_sync(0xffffffffu, f.cost, o);
    }
    return f;
}

// returns delta (uniform over the calling warp; all 32 lanes must call)
__device__ __forceinline__ float dirichlet_fit(float y, int lane, const DetectParams &p) {
    const bool active = lane < 7;
    float tab[4];
    const int li = active ? lane : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) tab[q] = p.fit_tab[li][q];
    float A = __shfl_sync(0xffffffffu, y, 3);
    float d = 0.f;
    float lambda = 1e-4f;
    FitSums f = fit_eval(y, active, A, d, tab, p.fit_W, p.fit_WoverN, p.fit_invN);
    for (int it = 0; it < 40; ++it) {
        // (J^T J + lambda diag) step = J^T r
        const float a11 = f.jaa * (1.f + lambda), a22 = f.jdd * (1.f + lambda), a12 = f.jad;
        const float det = a11 * a22 - a12 * a12;
        if (!(fabsf(det) > 0.f)) break;
        const float dA = (a22 * f.jar - a12 * f.jdr) / det;
        const float dd = (a11 * f.jdr - a12 * f.jar) / det;
        const float An = A + dA, dn = d + dd;
        const FitSums fn = fit_eval(y, active, An, dn, tab, p.fit_W, p.fit_WoverN, p.fit_invN);
        if (fn.cost <= f.cost) {
            A = An;
            d = dn;
            f = fn;
            lambda *= 0.1f;
            if (fabsf(dd) < 1e-7f && fabsf(dA) <= 1e-7f * fabsf(A)) break;
        } else {
            if (fabsf(dd) < 1e-7f && fabsf(dA) <= 1e-7f * fabsf(A)) break;   // converged to rounding
            lambda = lambda * 10.f + 1e-6f;
            if (lambda > 1e10f) break;
        }
    }
    return d;
}

// ------------------------------------------------------------------ the kernel
template <int LOG2N, int T, bool GMEM>
__global__ void __launch_bounds__(T, (T >= 512 ? 1 : (T >= 256 ? 2 : (T >= 128 ? 4 : 8))))
detect_kernel(const __grid_constant__ DetectParams p) {
    using C = Cfg<LOG2N, T, GMEM>;
    constexpr int N = C::N, M = C::M, R2 = C::R2, R3 = C::R3, S = C::S;
    constexpr int I1 = C::I1, I2 = C::I2, I3 = C::I3;
    constexpr int LOG2M = ilog2(M), LOG2R3 = ilog2(R3), LOG2R2 = ilog2(R2), LOG2S = ilog2(S);
    constexpr uint32_t RAW_BYTES = 2u * N;

    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;

    // ---- shared memory carve-up
    size_t off = 0;
    float2 *buf;
    if (GMEM) {
        buf = p.scratch + (size_t)blockIdx.x * N;
    } else {
        buf = reinterpret_cast<float2 *>(smem);
        off += (size_t)N * 8;
    }
    unsigned char *raw_s = smem + off;
    off += 2 * (size_t)RAW_BYTES;
    float2 *tw2 = reinterpret_cast<float2 *>(smem + off);
    off += (size_t)M * 8;
    float2 *rho = reinterpret_cast<float2 *>(smem + off);
    off += 32 * 8;
    uint32_t *red = reinterpret_cast<uint32_t *>(smem + off);   // 64 words reduction scratch
    float *bc = reinterpret_cast<float *>(smem + off + 256);    // broadcast scratch (64 floats)
    off += 1024;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + off);  // 2 barriers

    // ---- one-time per-CTA setup: twiddles
    for (int idx = tid; idx < M; idx += T) {
        const int k2 = idx >> LOG2R3, n3 = idx & (R3 - 1);
        const int e = (n3 * k2) & (M - 1);
        tw2[idx] = cispi(-2.0f * (float)e / (float)M);          // W_M^{n3 k2}
    }
    float2 w1[I1], w4[I1];                                      // W_N^j and W_N^{4j}
#pragma unroll
    for (int i = 0; i < I1; ++i) {
        const int j = tid + T * i;
        w1[i] = cispi(-2.0f * (float)j / (float)N);
        w4[i] = cispi(-2.0f * (float)((4 * j) & (N - 1)) / (float)N);
    }
    const bool use_raw = (p.raw != nullptr);
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    uint32_t par0 = 0, par1 = 0;    // phase parity of the two tile barriers
    int stage = 0;
    if (use_raw && tid == 0 && (int)blockIdx.x < p.n_blocks) {
        mbar_expect_tx(&mbar[0], RAW_BYTES);
        tma_bulk_g2s(raw_s, p.raw + (size_t)blockIdx.x * RAW_BYTES, RAW_BYTES, &mbar[0]);
    }

    for (int blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x) {
        // ---- prefetch the next block's raw tile into the other stage, wait for ours
        const unsigned char *rawt = raw_s + (size_t)stage * RAW_BYTES;
        if (use_raw) {
            const int nxt = blk + gridDim.x;
            if (tid == 0 && nxt < p.n_blocks) {
                mbar_expect_tx(&mbar[stage ^ 1], RAW_BYTES);
                tma_bulk_g2s(raw_s + (size_t)(stage ^ 1) * RAW_BYTES, p.raw + (size_t)nxt * RAW_BYTES,
                             RAW_BYTES, &mbar[stage ^ 1]);
            }
            mbar_wait(&mbar[stage], stage ? par1 : par0);
            if (stage) par1 ^= 1; else par0 ^= 1;
        }
        const float2 *iqb = use_raw ? nullptr : p.iq + (size_t)blk * N;
        const int64_t bidx = p.block_idx ? p.block_idx[blk] : (int64_t)blk;

        // sample loader: rawconv (block_data.py:38-52) or complex64 passthrough
        auto load_x = [&](int n) -> float2 {
            if (use_raw) {
                const uchar2 b = reinterpret_cast<const uchar2 *>(rawt)[n];
                return make_float2(((float)b.x - 127.4f) * 0.0078125f, ((float)b.y - 127.4f) * 0.0078125f);
            } else {
                return __ldg(&iqb[n]);
            }
        };

        // FFT buffer accessors: swizzled shared memory, or L2-only global scratch
        auto bld = [&](int e) -> float2 {
            if constexpr (GMEM) return __ldcg(&buf[e]);
            else return buf[C::pos(e)];
        };
        auto bst = [&](int e, float2 v) {
            if constexpr (GMEM) __stcg(&buf[e], v);
            else buf[C::pos(e)] = v;
        };

        // ================= forward FFT passes (shared by FFT#1 and FFT#2) =================
        // pass 1: radix-32 over n1 (stride M), twiddle W_N^{j k1}, in-place store
        auto fwd_pass1 = [&](auto &&loader) {
#pragma unroll
            for (int i = 0; i < I1; ++i) {
                const int j = tid + T * i;
                float xr[32], xi[32];
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const float2 v = loader(n1 * M + j, n1, i);
                    xr[n1] = v.x;
                    xi[n1] = v.y;
                }
                fft_dif<32>(xr, xi);
                float2 cur[4];
                cur[0] = w1[i];
                cur[1] = cmul(w1[i], w1[i]);
                cur[2] = cmul(cur[1], w1[i]);
                cur[3] = w4[i];
                bst(j, make_float2(xr[0], xi[0]));
#pragma unroll
                for (int k1 = 1; k1 < 32; ++k1) {
                    const int r = brev(k1, 5);
                    if (k1 > 4) cur[(k1 - 1) & 3] = cmul(cur[(k1 - 1) & 3], w4[i]);
                    bst(k1 * M + j, cmul(make_float2(xr[r], xi[r]), cur[(k1 - 1) & 3]));
                }
            }
        };
        // pass 2: radix-R2 over n2 (stride R3) inside each k1 slab, twiddle W_M^{n3 k2}
        auto fwd_pass2 = [&]() {
            if (R2 == 1) return;
#pragma unroll
            for (int i = 0; i < I2; ++i) {
                const int w = tid + T * i;
                const int k1 = w >> LOG2R3, n3 = w & (R3 - 1);
                const int base = k1 * M + n3;
                float xr[R2], xi[R2];
#pragma unroll
                for (int n2 = 0; n2 < R2; ++n2) {
                    const float2 v = bld(base + n2 * R3);
                    xr[n2] = v.x;
                    xi[n2] = v.y;
                }
                fft_dif<R2>(xr, xi);
#pragma unroll
                for (int k2 = 0; k2 < R2; ++k2) {
                    const int r = brev(k2, LOG2R2);
                    float2 v = make_float2(xr[r], xi[r]);
                    if (k2 > 0) v = cmul(v, tw2[k2 * R3 + n3]);
                    bst(base + k2 * R3, v);
                }
            }
        };

        // ================= FFT #1 =================
        fwd_pass1([&](int n, int, int) { return load_x(n); });
        __syncthreads();
        fwd_pass2();
        if (R2 > 1) __syncthreads();

        // pass 3 + power spectrum (Signal.mag, signal_utils.py:99-107) + windowed arg-max
        float pw[I3][R3];
        float esum = 0.f, msum = 0.f;
        unsigned long long best = 0ull;
        const bool need_std_c = (p.c_std != 0.f);
#pragma unroll
        for (int i = 0; i < I3; ++i) {
            const int g = tid + T * i;
            float xr[R3], xi[R3];
#pragma unroll
            for (int n3 = 0; n3 < R3; ++n3) {
                const float2 v = bld(g * R3 + n3);
                xr[n3] = v.x;
                xi[n3] = v.y;
            }
            fft_dif<R3>(xr, xi);
            const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));     // k1 + 32 k2
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) {
                const int r = brev(k3, LOG2R3);
                const float pv = xr[r] * xr[r] + xi[r] * xi[r];
                pw[i][k3] = pv;
                esum += pv;
                if (need_std_c) msum += sqrtf(pv);
                const int k = kb + S * k3;
                const uint32_t rel = (uint32_t)(k - p.win_start) & (uint32_t)(N - 1);
                if (rel < (uint32_t)p.win_len) {
                    const unsigned long long c = pack_cand(pv, rel);
                    best = c > best ? c : best;
                }
                if (p.dbg_fft_mag) p.dbg_fft_mag[k] = sqrtf(pv);
            }
        }
        const RedOut ra = block_reduce<T>(esum, msum, best, red, tid);

        // ---- carrier decision in float32 (carrier_detect.py:99-115)
        const float peak_pw = __uint_as_float((uint32_t)(ra.best >> 32));
        const uint32_t peak_rel = 0xffffffffu - (uint32_t)ra.best;
        const int kpeak = (p.win_start + (int)peak_rel) & (N - 1);
        const float peak_mag = sqrtf(peak_pw);
        const float noise_pw_c = (ra.s0 - 2.f * (peak_mag * peak_mag)) / (float)(N - 1);
        const float noise_c = sqrtf(noise_pw_c);
        float var_c = 0.f;
        if (need_std_c) {
            const float mean = ra.s1 / (float)N;
            var_c = ra.s0 / (float)N - mean * mean;
            var_c = sqrtf(fmaxf(var_c, 0.f));
            var_c = var_c * var_c;
        }
        const float thr_c = sqrtf(p.c_const + p.c_snr * (noise_c * noise_c) + p.c_std * var_c);
        const bool carrier = peak_mag > thr_c;

        if (!carrier) {
            if (tid < p.n_templates) {
                thr_record rec;
                rec.block_idx = bidx;
                rec.soa = __longlong_as_double(0x7ff8000000000000ll);
                rec.carrier_bin = kpeak;
                rec.carrier_offset = 0.f;
                rec.carrier_energy = peak_mag;
                rec.carrier_noise = noise_c;
                rec.corr_sample = -1;
                rec.corr_offset = __int_as_float(0x7fc00000);
                rec.corr_energy = __int_as_float(0x7fc00000);
                rec.corr_noise = __int_as_float(0x7fc00000);
                rec.flags = 0u;
                rec.template_idx = tid;
                rec.signal_energy = ra.s0 / (float)N;
                rec.reserved = 0.f;
                p.out[(size_t)blk * p.n_templates + tid] = rec;
            }
            stage ^= 1;
            __syncthreads();     // raw tile reads done before the next prefetch overwrites it
            continue;
        }

        // ---- gather the 7 magnitudes around the peak for the Dirichlet fit
#pragma unroll
        for (int i = 0; i < I3; ++i) {
            const int g = tid + T * i;
            const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
            const uint32_t u = (uint32_t)(kb - kpeak + 3) & (uint32_t)(N - 1);
            const uint32_t lo = u & (uint32_t)(S - 1);
            if (lo < 7u) {
                const int k3s = (R3 - (int)(u >> LOG2S)) & (R3 - 1);
                float v = 0.f;
#pragma unroll
                for (int k3 = 0; k3 < R3; ++k3) v = (k3 == k3s) ? pw[i][k3] : v;
                bc[lo] = sqrtf(v);
            }
        }
        __syncthreads();
        if (tid < 32) {
            const float y = lane < 7 ? bc[lane] : 0.f;
            const float d = dirichlet_fit(y, lane, p);
            // mix phasors for the 32 radix-1 positions: rho[n1] = exp(-2 pi i (k+d) n1 / 32)
            const int e = (kpeak * lane) & 31;
            const float turns = -((float)e * 0.03125f) - d * ((float)lane * 0.03125f);
            rho[lane] = cispi(2.f * turns);
            if (lane == 0) bc[8] = d;
        }
        __syncthreads();
        const float delta = bc[8];

        // ================= mix + FFT #2 (carrier_sync.py:222-238) =================
        // x'[n] = x[n] exp(2 pi i shift (n/N - 1/2)), shift = -(k + delta); n = n1*M + j
        float2 ph0[I1];
#pragma unroll
        for (int i = 0; i < I1; ++i) {
            const int j = tid + T * i;
            const int e = (int)(((long long)kpeak * j) & (N - 1));
            float turns = -((float)e / (float)N) - delta * ((float)j / (float)N);
            turns += 0.5f * (float)(kpeak & 1) + 0.5f * delta;
            ph0[i] = cispi(2.f * turns);
        }
        fwd_pass1([&](int n, int n1, int i) {
            const float2 ph = cmul(ph0[i], rho[n1]);
            return cmul(load_x(n), ph);
        });
        __syncthreads();
        fwd_pass2();
        if (R2 > 1) __syncthreads();

        // pass 3 of FFT#2, energy of X', then per template: x conj(T)/N and inverse pass 3'
        float e2sum = 0.f;
        for (int tpl = 0; tpl < p.n_templates; ++tpl) {
            const float2 *tsp = p.tpl_spec + (size_t)tpl * N;
#pragma unroll
            for (int i = 0; i < I3; ++i) {
                const int g = tid + T * i;
                float xr[R3], xi[R3];
                if (tpl == 0) {
#pragma unroll
                    for (int n3 = 0; n3 < R3; ++n3) {
                        const float2 v = bld(g * R3 + n3);
                        xr[n3] = v.x;
                        xi[n3] = v.y;
                    }
                    fft_dif<R3>(xr, xi);
                    const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
#pragma unroll
                    for (int k3 = 0; k3 < R3; ++k3) {
                        const int r = brev(k3, LOG2R3);
                        e2sum += xr[r] * xr[r] + xi[r] * xi[r];
                        if (p.dbg_shifted_fft) p.dbg_shifted_fft[kb + S * k3] = make_float2(xr[r], xi[r]);
                        if (p.n_templates > 1)
                            p.xsave[(size_t)blockIdx.x * N + (size_t)(i * R3 + k3) * T + tid] =
                                make_float2(xr[r], xi[r]);
                    }
                } else {
#pragma unroll
                    for (int k3 = 0; k3 < R3; ++k3) {
                        const int r = brev(k3, LOG2R3);
                        const float2 v = p.xsave[(size_t)blockIdx.x * N + (size_t)(i * R3 + k3) * T + tid];
                        xr[r] = v.x;
                        xi[r] = v.y;
                    }
                }
                // multiply by conj(T)/N (soa_estimator.py:99) and run the inverse radix-R3 DFT
                float yr[R3], yi[R3];
#pragma unroll
                for (int k3 = 0; k3 < R3; ++k3) {
                    const int r = brev(k3, LOG2R3);
                    const float2 t = __ldg(&tsp[(size_t)(i * R3 + k3) * T + tid]);
                    const float2 v = cmul(make_float2(xr[r], xi[r]), t);
                    yr[k3] = v.x;
                    yi[k3] = v.y;
                }
                fft_dif<R3>(yi, yr);          // inverse: swapped roles
#pragma unroll
                for (int n3 = 0; n3 < R3; ++n3) {
                    const int r = brev(n3, LOG2R3);
                    bst(g * R3 + n3, make_float2(yr[r], yi[r]));
                }
            }
            __syncthreads();
            // inverse pass 2': conj twiddle on load, radix-R2 over k2
            if (R2 > 1) {
#pragma unroll
                for (int i = 0; i < I2; ++i) {
                    const int w = tid + T * i;
                    const int k1 = w >> LOG2R3, n3 = w & (R3 - 1);
                    const int base = k1 * M + n3;
                    float xr[R2], xi[R2];
#pragma unroll
                    for (int k2 = 0; k2 < R2; ++k2) {
                        float2 v = bld(base + k2 * R3);
                        if (k2 > 0) v = cmulc(v, tw2[k2 * R3 + n3]);
                        xr[k2] = v.x;
                        xi[k2] = v.y;
                    }
                    fft_dif<R2>(xi, xr);
#pragma unroll
                    for (int n2 = 0; n2 < R2; ++n2) {
                        const int r = brev(n2, LOG2R2);
                        bst(base + n2 * R3, make_float2(xr[r], xi[r]));
                    }
                }
                __syncthreads();
            }
            // inverse pass 1': conj twiddle on load, radix-32 over k1 -> c[n1*M + j]
            float cp[I1][32];
            float c1sum = 0.f, c2sum = 0.f;
            unsigned long long cbest = 0ull;
            const bool need_std_k = (p.k_std != 0.f);
#pragma unroll
            for (int i = 0; i < I1; ++i) {
                const int j = tid + T * i;
                float xr[32], xi[32];
                float2 cur[4];
                cur[0] = w1[i];
                cur[1] = cmul(w1[i], w1[i]);
                cur[2] = cmul(cur[1], w1[i]);
                cur[3] = w4[i];
                {
                    const float2 v = bld(j);
                    xr[0] = v.x;
                    xi[0] = v.y;
                }
#pragma unroll
                for (int k1 = 1; k1 < 32; ++k1) {
                    if (k1 > 4) cur[(k1 - 1) & 3] = cmul(cur[(k1 - 1) & 3], w4[i]);
                    const float2 v = cmulc(bld(k1 * M + j), cur[(k1 - 1) & 3]);
                    xr[k1] = v.x;
                    xi[k1] = v.y;
                }
                fft_dif<32>(xi, xr);
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const int r = brev(n1, 5);
                    const int n = n1 * M + j;
                    const float pv = xr[r] * xr[r] + xi[r] * xi[r];
                    cp[i][n1] = pv;
                    if (n >= p.corr_start && n < p.corr_stop) {
                        const unsigned long long c = pack_cand(pv, (uint32_t)n);
                        cbest = c > cbest ? c : cbest;
                    }
                    if (need_std_k && n < p.corr_len) {
                        c1sum += sqrtf(pv);
                        c2sum += pv;
                    }
                    if (p.dbg_corr && tpl == 0 && n < p.corr_len) p.dbg_corr[n] = make_float2(xr[r], xi[r]);
                }
            }
            // reduce: (energy of X' | sum|c|), sum|c|^2, arg-max
            const RedOut rb = block_reduce<T>(need_std_k ? c1sum : e2sum, need_std_k ? c2sum : 0.f, cbest, red, tid);
            float e2tot;
            if (need_std_k) {
                // need the X' energy as well: second (cheap) reduction
                const RedOut rc = block_reduce<T>(e2sum, 0.f, 0ull, red, tid);
                e2tot = rc.s0;
            } else {
                e2tot = rb.s0;
            }
            const float peak_cp = __uint_as_float((uint32_t)(rb.best >> 32));
            const int s = (int)(0xffffffffu - (uint32_t)rb.best);
            // neighbours of the peak for the Gaussian interpolation
#pragma unroll
            for (int i = 0; i < I1; ++i) {
                const int j = tid + T * i;
#pragma unroll
                for (int dd = -1; dd <= 1; dd += 2) {
                    const int nt = s + dd;
                    if (nt >= 0 && nt < N && (nt & (M - 1)) == j) {
                        const int n1s = nt >> LOG2M;
                        float v = 0.f;
#pragma unroll
                        for (int n1 = 0; n1 < 32; ++n1) v = (n1 == n1s) ? cp[i][n1] : v;
                        bc[16 + (dd + 1)] = v;
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                // float64 scalar tail (soa_estimator.py:78-134,159-170)
                const double peak_mag_k = sqrt((double)peak_cp);
                const double sig_energy = (double)e2tot / (double)N;
                const double noise_pw = (sig_energy * (double)p.tpl_energy[tpl] - (double)peak_cp) / (double)N;
                const double noise_k = sqrt(noise_pw);                     // NaN if negative
                double var_k = 0.0;
                if (need_std_k) {
                    const double mean = (double)rb.s0 / (double)p.corr_len;
                    var_k = (double)rb.s1 / (double)p.corr_len - mean * mean;
                    if (var_k < 0.0) var_k = 0.0;
                }
                const double thr_k = sqrt((double)p.k_const + (double)p.k_snr * (noise_k * noise_k) +
                                          (double)p.k_std * var_k);
                const bool detected = peak_mag_k > thr_k;
                double offset = 0.0;
                if (detected && s > 0 && s < p.corr_len - 1) {
                    const double pa = (double)bc[16], pc = (double)bc[18], pb = (double)peak_cp;
                    // a,b,c = ln|c|: offset = 0.5 (c-a) / (2b-a-c), on magnitudes = sqrt(power)
                    const double num = 0.5 * log(pc / pa);
                    const double den = 0.5 * log(pb * pb / (pa * pc));
                    offset = 0.5 * num / den;
                    if (!(offset == offset)) offset = __longlong_as_double(0x7ff8000000000000ll);
                    offset = offset < -0.6 ? -0.6 : (offset > 0.6 ? 0.6 : offset);
                }
                thr_record rec;
                rec.block_idx = bidx;
                rec.soa = (double)p.new_len * (double)bidx + (double)s + offset;
                rec.carrier_bin = kpeak;
                rec.carrier_offset = delta;
                rec.carrier_energy = peak_mag;
                rec.carrier_noise = noise_c;
                rec.corr_sample = s;
                rec.corr_offset = (float)offset;
                rec.corr_energy = (float)peak_mag_k;
                rec.corr_noise = (float)noise_k;
                rec.flags = THR_FLAG_CARRIER_DETECTED | (detected ? THR_FLAG_CORR_DETECTED : 0u);
                rec.template_idx = tpl;
                rec.signal_energy = (float)sig_energy;
                rec.reserved = 0.f;
                p.out[(size_t)blk * p.n_templates + tpl] = rec;
            }
            __syncthreads();
        }
        stage ^= 1;
    }
}

}  // namespace thr
