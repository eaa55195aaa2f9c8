// detect_kernel_2x.cuh -- fused detect kernel for block_len = 32768 (sm_100a).
//
// A 32768-point complex buffer (256 KiB) does not fit next to anything else in the 227 KiB of shared
// memory of an SM, so the block is transformed as TWO interleaved 16384-point transforms that take
// turns in the one 136 KiB buffer the N = 16384 kernel uses, joined by a radix-2 step in registers:
//
//     E = FFT_F(x[2m]),  O = FFT_F(x[2m+1]),  F = 16384,  w_k = W_{2F}^k
//     X[k] = E[k] + w_k O[k],   X[k+F] = E[k] - w_k O[k]                         (decimation in time)
//     c[2m]   = IFFT_F(A)[m],   A[k] = Y[k] + Y[k+F]
//     c[2m+1] = IFFT_F(B)[m],   B[k] = (Y[k] - Y[k+F]) conj(w_k)                 (decimation in frequency)
//
// with Y = X' x conj(T)/N.  After pass 3 every thread holds its 32 spectrum values in registers, so the
// half that has to wait (E while O is transformed, B while A is inverted) is parked in a per-CTA,
// L2-resident scratch area with thread-private coalesced st.cg / ld.cg (128 KiB each way); nothing else
// leaves the SM.  |c|^2 of every lag is also written there so that the neighbours of the peak (which
// belong to the other half-transform) can be fetched without keeping 64 more registers alive.
//
// Scope: every configuration of the Python path's semantics at this block length.  FFT#1 is pruned ("zoom") when the
// carrier window (+-3 bins) is at most 128 bins wide (moved to bin 0 by an integer pre-shift if necessary) and there is no
// carrier stddev threshold term; otherwise both half transforms of FFT#1 are computed in full and joined the same way as
// FFT#2, with |X|^2 of all 32768 bins parked in the scratch area for the arg-max key and the 7 fit magnitudes.  With
// several templates O' and the odd-lag spectrum get parking areas of their own (DetectParams::xsave), so that E' and O'
// survive the template loop.  The generic global-scratch variant of detect_kernel remains for debug launches.
//
// Same semantics, mailboxes, service-warp pipeline and record format as detect_kernel (see there for the
// reference citations of each stage).
#pragma once

#include <type_traits>

#include "detect_kernel.cuh"

#ifndef THR_2X_SMEM_B
#define THR_2X_SMEM_B 1     // stage B reads the raw block from the shared-memory stage too (a second TMA fetch of the tile, an
#endif                      // L2 hit hidden behind the end of stage A) instead of 64 ld.global.nc.u16 per thread and block

namespace thr {

struct Cfg2x {
    static constexpr int NB = 32768;                 // block length
    static constexpr int F = 16384;                  // transform length of each half
    static constexpr int T = 512;
    static constexpr int M = F / 32;                 // 512
    static constexpr int R2 = 32, R3 = 16, S = 32 * R2;
    static constexpr int LAUNCH_THREADS = T + 128;
    static constexpr size_t BUF_BYTES = (size_t)(F / 16) * 136;
    static constexpr int MAX_TPL = 32;               // templates per detector (tail mailbox size, as Cfg::MAX_TPL)
    static constexpr size_t smem_bytes() {
        return BUF_BYTES + (size_t)(2 * NB) + (size_t)M * 8 + 2 * 320 + 2 * 32 + 2 * MAX_TPL * 32 + 256
               + 2 * 128 * 8 + 512 + 256 + 64 + sizeof(lm::Rows);
    }
    using Half = Cfg<14, 512, false>;                // geometry of one half (pass-3 item order)
};

// MULTI = false: exactly one template (no template loop, B parked over E'), as detect_kernel's MULTI
template <bool MULTI>
__global__ void __launch_bounds__(Cfg2x::LAUNCH_THREADS, 1)
detect2x_kernel(const __grid_constant__ DetectParams p) {
    using C = Cfg2x;
    using H = Cfg2x::Half;
    constexpr int NB = C::NB, F = C::F, T = C::T, M = C::M, R2 = C::R2, R3 = C::R3;
    constexpr int LOG2M = 9, LOG2R3 = 4, LOG2R2 = 5;
    constexpr uint32_t RAW_BYTES = 2u * NB;
    constexpr int NTHREADS = T + 32;
    constexpr uint32_t A1_STEP = (uint32_t)(M / 16) * 136u, A2_STEP = 136u;
    constexpr int S = 32 * R2;                       // bin stride of the pass-3 outputs

    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;

    size_t off = 0;
    unsigned char *bufc = smem;
    off += C::BUF_BYTES;
    unsigned char *raw_s = smem + off;                                   // ONE raw stage (64 KiB)
    off += RAW_BYTES;
    float2 *tw2 = reinterpret_cast<float2 *>(smem + off);
    off += (size_t)M * 8;
    FitSlot *fitslot = reinterpret_cast<FitSlot *>(smem + off);          // [2]
    off += 2 * sizeof(FitSlot);
    TailHdr *tailhdr = reinterpret_cast<TailHdr *>(smem + off);          // [2]
    off += 2 * sizeof(TailHdr);
    constexpr int TPL_SLOTS = MULTI ? C::MAX_TPL : 1;                    // tail mailbox entries per block
    TailSlot *tailslot = reinterpret_cast<TailSlot *>(smem + off);       // [2][TPL_SLOTS]
    off += 2 * (size_t)TPL_SLOTS * sizeof(TailSlot);
    uint32_t *red = reinterpret_cast<uint32_t *>(smem + off);
    off += 256;
    float2 *zc = reinterpret_cast<float2 *>(smem + off);                 // [2][128] pruned spectra of E, O
    off += 2 * 128 * 8;
    float *zpow = reinterpret_cast<float *>(smem + off);                 // |X[b0 + k]|^2, k < 128
    off += 512;
    float2 *zrho = reinterpret_cast<float2 *>(smem + off);               // W_32^{b0 n1}: row phasors of the zoom pre-shift
    off += 256;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + off);
    off += 64;
    lm::Rows &fitrows = *reinterpret_cast<lm::Rows *>(smem + off);       // row workspace of the Dirichlet fit

#define use_raw (p.raw != nullptr)
#define need_std_k (p.k_std != 0.f)
    auto has_block = [&](int i) -> bool { return (int)blockIdx.x + i * (int)gridDim.x < p.n_blocks; };
    float2 *scrE = p.scratch + (size_t)blockIdx.x * NB;                  // F float2: parked half spectrum
    float *cps = reinterpret_cast<float *>(scrE + F);                    // [2][F] floats: |c|^2 per lag parity
    // several templates: O' (pass-3 outputs of the odd half) and the odd-lag spectrum B get their own parking areas, so
    // that E' and O' survive the template loop; with one template B simply overwrites E'
    const int n_tpl = MULTI ? p.n_templates : 1;
    float2 *scrO = MULTI ? p.xsave + (size_t)blockIdx.x * NB : nullptr;
    float2 *scrB = MULTI ? scrO + F : scrE;

    asm volatile("griddepcontrol.launch_dependents;");
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_fence_init();
    }
    for (int idx = tid; idx < M; idx += C::LAUNCH_THREADS) {
        const int k2 = idx >> LOG2R3, n3 = idx & (R3 - 1);
        tw2[idx] = cispi(-2.0f * (float)((n3 * k2) & (M - 1)) / (float)M);
    }
    if (tid < 32) zrho[tid] = cispi(-2.0f * (float)((p.zoom_base * tid) & 31) * 0.03125f);
    __syncthreads();

    // ---- serial work (service warp): Dirichlet fit + mix phasor table, scalar tail + record
    auto do_fit = [&](int q) {
        FitSlot &fs = fitslot[q];
        if (fs.carrier) {
            const float y = (lane & 7) < 7 ? fs.mags[lane & 7] : 0.f;
            const float d = dirichlet_fit(y, lane, p, fitrows);
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
                const TailSlot &ts = tailslot[q * TPL_SLOTS + tpl];
                const float peak_mag_k = sqrtf(ts.peak_cp);
                const float noise_pw = (h.sig_energy1 * p.tpl_energy[tpl] - ts.peak_cp) / (float)NB;
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
                    const float num = logf(ts.pc / ts.pa);
                    const float den = logf((ts.peak_cp / ts.pa) * (ts.peak_cp / ts.pc));
                    offset = fminf(fmaxf(0.5f * num / den, -0.6f), 0.6f);
                }
                rec.soa = (double)p.new_len * (double)bidx + (double)ts.s + (double)offset;
                rec.carrier_offset = h.delta;
                rec.corr_sample = ts.s;
                rec.corr_offset = offset;
                rec.corr_energy = peak_mag_k;
                rec.corr_noise = noise_k;
                rec.flags = THR_FLAG_CARRIER_DETECTED | (detected ? THR_FLAG_CORR_DETECTED : 0u);
            }
            p.out[(size_t)blk * n_tpl + tpl] = rec;
        }
    };

    if (tid >= T) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (tid >= T + 32) return;
        for (int i = -1; has_block(i < 0 ? 0 : i); ++i) {
            if (has_block(i + 1)) {
                const int q = (i + 1) & 1;
                bar_sync(BAR_FITREQ + q, NTHREADS);
                do_fit(q);
                bar_arrive(BAR_FITDONE + q, NTHREADS);
            }
            if (i >= 0) {
                const int q = i & 1;
                bar_sync(BAR_TAILREQ + q, NTHREADS);
                do_tail(i, q);
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");

    // ---- FFT buffer (shared memory, rows of 16 complex values padded to 17)
    auto ld8 = [&](uint32_t byte_off) -> float2 { return *reinterpret_cast<const float2 *>(bufc + byte_off); };
    auto st8 = [&](uint32_t byte_off, float2 v) { *reinterpret_cast<float2 *>(bufc + byte_off) = v; };
    const uint32_t a1b = (uint32_t)(tid >> 4) * 136u + (uint32_t)(tid & 15) * 8u;          // pass 1: item j = tid
    const uint32_t a2b = (uint32_t)(tid >> 4) * A1_STEP + (uint32_t)(tid & 15) * 8u;       // pass 2: (k1, n3)
    const int p2_n3 = tid & (R3 - 1);

    const float2 w1 = cispi(-2.0f * (float)tid / (float)F);                                // W_F^j
    const float2 w4 = cispi(-2.0f * (float)((4 * tid) & (F - 1)) / (float)F);              // W_F^{4j}

    auto issue_tile = [&](int i) {
        const int blk = (int)blockIdx.x + i * (int)gridDim.x;
        mbar_expect_tx(&mbar[0], RAW_BYTES);
        tma_bulk_g2s(raw_s, p.raw + (size_t)blk * (size_t)p.raw_stride, RAW_BYTES, &mbar[0]);
    };
    if (use_raw && tid == 0 && has_block(0)) issue_tile(0);
    uint32_t par = 0;
    // The one raw stage serves both stages of the pipeline (THR_2X_SMEM_B): once pass 1 of A(ia) has consumed tile ia, the
    // stage is refilled with tile ia-1 for B(ia-1) if that block has a carrier (decided one iteration ago, written by this
    // very thread), else with tile ia+1 for A(ia+1); B refills it with tile ia+1 after its own pass 1.  Every refill comes
    // behind a BAR_MAIN that follows all threads' wait for the previous fill (mbarrier phase safety).
    auto refill_after_a = [&](int ia) {              // thread 0, after the pass-1 barrier of A(ia), second half
        if (THR_2X_SMEM_B && ia >= 1 && fitslot[(ia - 1) & 1].carrier) issue_tile(ia - 1);
        else if (has_block(ia + 1)) issue_tile(ia + 1);
    };

    // forward pass 1 of half h of block i: samples x[2m + h], m = n1*M + j (from the shared-memory raw
    // stage in both stages; complex64 input: from global memory) -> radix-32 over n1 -> twiddle W_F^{j k1}
    //   SHIFT : multiply by the phasor rho[n1] * ph0 (mix of stage B, or the integer pre-shift of the zoom band)
    //   STAGEB: FFT#2 (the tile was fetched a second time, see refill_after_a; no energy sum); otherwise FFT#1
    auto pass1 = [&](auto shift_c, auto stageb_c, int i, int h, float2 ph0, const float2 *rho, float &energy) {
        constexpr bool shift = decltype(shift_c)::value, stageb = decltype(stageb_c)::value;
        const int blk = (int)blockIdx.x + i * (int)gridDim.x;
        float2 x[32];
        if (use_raw) {
            if constexpr (!stageb) {
                const uint16_t *rawt = reinterpret_cast<const uint16_t *>(raw_s);
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) x[brev(n1, 5)] = rawconv(rawt[2 * (n1 * M + tid) + h]);
            } else if (THR_2X_SMEM_B) {
                const uint16_t *rawt = reinterpret_cast<const uint16_t *>(raw_s);
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) x[brev(n1, 5)] = rawconv(rawt[2 * (n1 * M + tid) + h]);
            } else {
                const uint16_t *rawg = reinterpret_cast<const uint16_t *>(p.raw + (size_t)blk * (size_t)p.raw_stride);
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) x[brev(n1, 5)] = rawconv(__ldg(&rawg[2 * (n1 * M + tid) + h]));
            }
        } else {
            const float2 *iqb = p.iq + (size_t)blk * NB;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) x[brev(n1, 5)] = __ldg(&iqb[2 * (n1 * M + tid) + h]);
        }
        if constexpr (!stageb) {
            float2 e2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) e2 = __ffma2_rn(x[n1], x[n1], e2);
            energy += e2.x + e2.y;
        }
        if constexpr (shift) {
            const float4 *rho4 = reinterpret_cast<const float4 *>(rho);
#pragma unroll
            for (int n1 = 0; n1 < 32; n1 += 2) {
                const float4 r = rho4[n1 >> 1];
                if (n1 > 0) x[brev(n1, 5)] = cmul(x[brev(n1, 5)], make_float2(r.x, r.y));
                x[brev(n1 + 1, 5)] = cmul(x[brev(n1 + 1, 5)], make_float2(r.z, r.w));
            }
        }
        fft_dit<32, false>(x);
        float2 ws = w1, ws4 = w4;
        asm volatile("" : "+f"(ws.x), "+f"(ws.y), "+f"(ws4.x), "+f"(ws4.y));
        float2 cur[4];
        cur[0] = ws;
        cur[1] = cmul(ws, ws);
        cur[2] = cmul(cur[1], ws);
        cur[3] = ws4;
        if constexpr (shift) {
#pragma unroll
            for (int c = 0; c < 4; ++c) cur[c] = cmul(cur[c], ph0);
            st8(a1b, cmul(x[0], ph0));
        } else {
            st8(a1b, x[0]);
        }
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) {
            if (k1 > 4) cur[(k1 - 1) & 3] = cmul(cur[(k1 - 1) & 3], ws4);
            st8(a1b + (uint32_t)k1 * A1_STEP, cmul(x[k1], cur[(k1 - 1) & 3]));
        }
    };
    // forward pass 2 (full): radix-32 over n2 inside each k1 slab, twiddle W_M^{n3 k2}
    auto pass2 = [&]() {
        float2 x[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) x[brev(n2, LOG2R2)] = ld8(a2b + (uint32_t)n2 * A2_STEP);
        fft_dit<R2, false>(x);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
            float2 v = x[k2];
            if (k2 > 0) v = cmul(v, tw2[k2 * R3 + p2_n3]);
            st8(a2b + (uint32_t)k2 * A2_STEP, v);
        }
    };
    // forward pass 2 pruned to the outputs k2 = 0..3 (bins < 128), see detect_kernel
    auto pass2_pruned = [&]() {
        float2 b0, b1, b2, b3;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 a0 = ld8(a2b + (uint32_t)(r) * A2_STEP);
            const float2 a1 = ld8(a2b + (uint32_t)(8 + r) * A2_STEP);
            const float2 a2 = ld8(a2b + (uint32_t)(16 + r) * A2_STEP);
            const float2 a3 = ld8(a2b + (uint32_t)(24 + r) * A2_STEP);
            const float2 s0 = f2add(a0, a2), s1 = f2sub(a0, a2);
            const float2 s2 = f2add(a1, a3), s3 = f2sub(a1, a3);
            const float2 c0 = f2add(s0, s2), c2 = f2sub(s0, s2);
            const float2 c1 = f2add(s1, rot_mj(s3)), c3 = f2add(s1, rot_pj(s3));
            if (r == 0) {
                b0 = c0; b1 = c1; b2 = c2; b3 = c3;
            } else {
                b0 = f2add(b0, c0);
                if (r == 1) { b1 = fma_tw32<1>(b1, c1); b2 = fma_tw32<2>(b2, c2); b3 = fma_tw32<3>(b3, c3); }
                if (r == 2) { b1 = fma_tw32<2>(b1, c1); b2 = fma_tw32<4>(b2, c2); b3 = fma_tw32<6>(b3, c3); }
                if (r == 3) { b1 = fma_tw32<3>(b1, c1); b2 = fma_tw32<6>(b2, c2); b3 = fma_tw32<9>(b3, c3); }
                if (r == 4) { b1 = fma_tw32<4>(b1, c1); b2 = fma_tw32<8>(b2, c2); b3 = fma_tw32<12>(b3, c3); }
                if (r == 5) { b1 = fma_tw32<5>(b1, c1); b2 = fma_tw32<10>(b2, c2); b3 = fma_tw32<15>(b3, c3); }
                if (r == 6) { b1 = fma_tw32<6>(b1, c1); b2 = fma_tw32<12>(b2, c2); b3 = fma_tw32<18>(b3, c3); }
                if (r == 7) { b1 = fma_tw32<7>(b1, c1); b2 = fma_tw32<14>(b2, c2); b3 = fma_tw32<21>(b3, c3); }
            }
        }
        st8(a2b, b0);
        st8(a2b + 1u * A2_STEP, cmul(b1, tw2[1 * R3 + p2_n3]));
        st8(a2b + 2u * A2_STEP, cmul(b2, tw2[2 * R3 + p2_n3]));
        st8(a2b + 3u * A2_STEP, cmul(b3, tw2[3 * R3 + p2_n3]));
    };
    // inverse pass 2': conj twiddle on load, radix-32 over k2
    auto pass2_inv = [&]() {
        float2 x[R2];
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
            float2 v = ld8(a2b + (uint32_t)k2 * A2_STEP);
            if (k2 > 0) v = cmulc(v, tw2[k2 * R3 + p2_n3]);
            x[brev(k2, LOG2R2)] = v;
        }
        fft_dit<R2, true>(x);
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) st8(a2b + (uint32_t)n2 * A2_STEP, x[n2]);
    };

    for (int i = -1; has_block(i < 0 ? 0 : i); ++i) {
        // ================================================================= A(i+1): pruned FFT #1 of both halves
        if (has_block(i + 1)) {
            const int ia = i + 1, q = ia & 1;
            if (use_raw) {
                mbar_wait(&mbar[0], par);
                par ^= 1;
            }
            float tenergy = 0.f;
            FitSlot &fs = fitslot[q];
            if (!p.zoom) {
                // ---- full FFT #1 (wide carrier window or a stddev threshold term): both halves are transformed
                // completely, E is parked like E' of stage B, the radix-2 join happens on this thread's pass-3 outputs of
                // O; |X|^2 of all NB bins goes to the (idle) |c|^2 area of the scratch so that the arg-max key and the
                // 7 fit magnitudes can be fetched without keeping 64 powers per thread in registers
                float *pws = cps;
                float bestv = 0.f, msum = 0.f;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    pass1(std::false_type{}, std::false_type{}, ia, h, make_float2(1.f, 0.f), nullptr, tenergy);
                    bar_sync(BAR_MAIN, T);
                    // the raw stage is not read again in stage A: fetch the next block's tile into it
                    if (h == 1 && use_raw && tid == 0) refill_after_a(ia);
                    pass2();
                    __syncwarp();
#pragma unroll 1
                    for (int it = 0; it < 2; ++it) {
                        const int g = H::p3_item(tid, it);
                        const uint32_t ab = (uint32_t)g * 136u;
                        float2 x[R3];
#pragma unroll
                        for (int n3 = 0; n3 < R3; ++n3) x[brev(n3, LOG2R3)] = ld8(ab + (uint32_t)n3 * 8u);
                        fft_dit<R3, false>(x);
                        float2 *park = &scrE[(size_t)(it * R3) * T + tid];
                        if (h == 0) {                                                    // E: parked
#pragma unroll
                            for (int k3 = 0; k3 < R3; ++k3) __stcg(&park[(size_t)k3 * T], x[k3]);
                        } else {                                                         // O: join with E
                            const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));          // bin k = kb + S k3 < F
                            const float2 wb = cispi(-2.0f * (float)kb / (float)NB);      // W_NB^kb
                            const uint32_t relb = (uint32_t)(kb - p.win_start) & (uint32_t)(NB - 1);
#pragma unroll
                            for (int k3 = 0; k3 < R3; ++k3) {
                                const float2 w = (k3 == 0) ? wb : mul_tw<false>(wb, cos32(k3), sin32(k3));
                                const float2 t = cmul(x[k3], w);
                                // (loading all 16 parked values ahead of this loop, or ahead of the butterflies, costs
                                // registers the rolled loops around it do not have: measured 1.04 / 1.12 instead of 0.96 ms)
                                const float2 ev = __ldcg(&park[(size_t)k3 * T]);
                                const float2 lo = f2add(ev, t), hi = f2sub(ev, t);       // X[k], X[k + F]
                                const float plo = lo.x * lo.x + lo.y * lo.y, phi = hi.x * hi.x + hi.y * hi.y;
                                const int k = kb + S * k3;
                                __stcg(&pws[k], plo);
                                __stcg(&pws[k + F], phi);
                                if (p.c_std != 0.f) msum += sqrtf(plo) + sqrtf(phi);
                                const uint32_t rlo = (relb + (uint32_t)(S * k3)) & (uint32_t)(NB - 1);
                                const uint32_t rhi = (rlo + (uint32_t)F) & (uint32_t)(NB - 1);
                                if (rlo < (uint32_t)p.win_len) bestv = fmaxf(bestv, plo);
                                if (rhi < (uint32_t)p.win_len) bestv = fmaxf(bestv, phi);
                            }
                        }
                    }
                    if (h == 0) bar_sync(BAR_MAIN, T);      // every warp is done reading the buffer
                }
                // first maximum in window order (np.argmax over the wrapped window, carrier_detect.py:138-149); the
                // thread that holds the maximum re-reads its own powers from the scratch
                const ArgOut ra = main_argmax<T, true>(__float_as_uint(bestv), tenergy, msum, red, tid, [&](uint32_t gb) {
                    uint32_t key = 0xffffffffu;
#pragma unroll 1
                    for (int it = 0; it < 2; ++it) {
                        const int g = H::p3_item(tid, it);
                        const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));
                        const uint32_t relb = (uint32_t)(kb - p.win_start) & (uint32_t)(NB - 1);
#pragma unroll 1
                        for (int k3 = 0; k3 < R3; ++k3) {
                            const int k = kb + S * k3;
                            const uint32_t rlo = (relb + (uint32_t)(S * k3)) & (uint32_t)(NB - 1);
                            const uint32_t rhi = (rlo + (uint32_t)F) & (uint32_t)(NB - 1);
                            if (rlo < (uint32_t)p.win_len && __float_as_uint(__ldcg(&pws[k])) == gb) key = min(key, rlo);
                            if (rhi < (uint32_t)p.win_len && __float_as_uint(__ldcg(&pws[k + F])) == gb) key = min(key, rhi);
                        }
                    }
                    return key;
                });
                // carrier decision in float32 (carrier_detect.py:99-115); Parseval: sum |X|^2 = NB sum |x|^2
                const float s0 = ra.s0 * (float)NB;
                const float peak_mag = sqrtf(__uint_as_float(ra.vbits));
                const int kpeak = (p.win_start + (int)ra.key) & (NB - 1);
                const float noise_c = sqrtf((s0 - 2.f * (peak_mag * peak_mag)) / (float)(NB - 1));
                float var_c = 0.f;
                if (p.c_std != 0.f) {
                    const float mean = ra.s1 / (float)NB;
                    var_c = fmaxf(s0 / (float)NB - mean * mean, 0.f);
                }
                const bool carrier = peak_mag > sqrtf(p.c_const + p.c_snr * (noise_c * noise_c) + p.c_std * var_c);
                if (tid == 0) {
                    fs.kpeak = kpeak;
                    fs.carrier = carrier ? 1 : 0;
                    fs.peak_mag = peak_mag;
                    fs.noise_c = noise_c;
                    fs.sig_energy1 = s0 / (float)NB;
                    fs.delta = 0.f;
                }
                // (every power was written before the two barriers of the arg-max: visible to these threads)
                if (carrier && tid < 7) fs.mags[tid] = sqrtf(__ldcg(&pws[(kpeak - 3 + tid) & (NB - 1)]));
                bar_arrive(BAR_FITREQ + q, NTHREADS);
            } else {
            const bool shiftA = (p.zoom_base != 0);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                if (shiftA) {     // zoom band not at bin 0: x[n] exp(-2 pi i b0 n / NB), n = 2 (n1 M + j) + h
                    const int e = (int)(((long long)p.zoom_base * tid) & (F - 1));
                    const float turns = -((float)e / (float)F) - (float)h * ((float)p.zoom_base / (float)NB);
                    pass1(std::true_type{}, std::false_type{}, ia, h, cispi(2.f * turns), zrho, tenergy);
                } else {
                    pass1(std::false_type{}, std::false_type{}, ia, h, make_float2(1.f, 0.f), nullptr, tenergy);
                }
                bar_sync(BAR_MAIN, T);
                // the raw stage is not read again in stage A: fetch the next block's tile into it
                if (h == 1 && use_raw && tid == 0) refill_after_a(ia);
                pass2_pruned();
                __syncwarp();
                {   // pruned pass 3: warp w owns slabs k1 = 2w, 2w+1: 8 bins x 4 lanes (4 terms each)
                    const int k1 = 2 * (tid >> 5) + (lane >> 4), k2 = (lane >> 2) & 3, sub = lane & 3;
                    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        acc = f2add(acc, ld8((uint32_t)k1 * A1_STEP + (uint32_t)(sub * 4 + t) * 8u + (uint32_t)k2 * A2_STEP));
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
                    if (sub == 0) zc[h * 128 + k1 + 32 * k2] = acc;
                }
                if (h == 1) {
                    const float wsum = warp_sum(tenergy);
                    if (lane == 0) red[16 + (tid >> 5)] = __float_as_uint(wsum);
                }
                bar_sync(BAR_MAIN, T);      // buffer free for the next pass 1; zc / energy partials visible
            }
            // radix-2 join for the 128 bins: X[k] = E[k] + W_NB^k O[k]
            if (tid < 128) {
                const float2 xk = f2add(zc[tid], cmul(zc[128 + tid], cispi(-2.0f * (float)tid / (float)NB)));
                zpow[tid] = xk.x * xk.x + xk.y * xk.y;
            }
            bar_sync(BAR_MAIN, T);
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
            key = __reduce_min_sync(0xffffffffu, key);
            float esum = 0.f;
#pragma unroll
            for (int w = 0; w < T / 32; ++w) esum += __uint_as_float(red[16 + w]);
            // carrier decision in float32 (carrier_detect.py:99-115); Parseval: sum |X|^2 = NB sum |x|^2
            const float s0 = esum * (float)NB;
            const float peak_mag = sqrtf(__uint_as_float(gb));
            const int kpeak = (p.win_start + (int)key) & (NB - 1);
            const float noise_c = sqrtf((s0 - 2.f * (peak_mag * peak_mag)) / (float)(NB - 1));
            const bool carrier = peak_mag > sqrtf(p.c_const + p.c_snr * (noise_c * noise_c));
            if (tid == 0) {
                fs.kpeak = kpeak;
                fs.carrier = carrier ? 1 : 0;
                fs.peak_mag = peak_mag;
                fs.noise_c = noise_c;
                fs.sig_energy1 = s0 / (float)NB;
                fs.delta = 0.f;
            }
            if (carrier && tid < 7) fs.mags[tid] = sqrtf(zpow[p.zoom_w0 + (int)key - 3 + tid]);
            bar_arrive(BAR_FITREQ + q, NTHREADS);
            }   // zoom
        }

        // ================================================================= B(i): mix, FFT #2, correlation
        if (i >= 0) {
            const int q = i & 1;
            bar_sync(BAR_FITDONE + q, NTHREADS);
            const FitSlot &fs = fitslot[q];
            const int kpeak = fs.kpeak;
            const bool carrier = fs.carrier != 0;
            if (tid == 0) {
                TailHdr &hd = tailhdr[q];
                hd.kpeak = kpeak;
                hd.carrier = fs.carrier;
                hd.peak_mag = fs.peak_mag;
                hd.noise_c = fs.noise_c;
                hd.sig_energy1 = fs.sig_energy1;
                hd.delta = fs.delta;
            }
            if (!carrier) {
                bar_arrive(BAR_TAILREQ + q, NTHREADS);
                continue;
            }
            const float delta = fs.delta;
            // x'[n] = x[n] exp(-2 pi i s (n/NB - 1/2)), s = k + delta, n = 2m + h, m = n1*M + j:
            //   row phasor rho[n1] (service warp), per-thread phasor exp(-2 pi i s (j/F + h/NB - 1/2))
            const int e0 = (int)(((long long)kpeak * tid) & (F - 1));
            const float turns0 = -((float)e0 / (float)F) - delta * ((float)tid / (float)F)
                                 + 0.5f * (float)(kpeak & 1) + 0.5f * delta;
            const float turns_h = -((float)kpeak + delta) / (float)NB;
            float unused_energy = 0.f;
            if (THR_2X_SMEM_B && use_raw) {
                // tile i again (an L2 hit), requested behind pass 1 of A(i+1); the last block of a CTA has no A(i+1)
                if (!has_block(i + 1) && tid == 0) issue_tile(i);
                mbar_wait(&mbar[0], par);
                par ^= 1;
            }

            // ---- half 0: E' = FFT_F(x'[2m]) -> parked
            pass1(std::true_type{}, std::true_type{}, i, 0, cispi(2.f * turns0), fs.rho, unused_energy);
            bar_sync(BAR_MAIN, T);
            pass2();
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const uint32_t ab = (uint32_t)H::p3_item(tid, it) * 136u;
                float2 x[R3];
#pragma unroll
                for (int n3 = 0; n3 < R3; ++n3) x[brev(n3, LOG2R3)] = ld8(ab + (uint32_t)n3 * 8u);
                fft_dit<R3, false>(x);
#pragma unroll
                for (int k3 = 0; k3 < R3; ++k3) __stcg(&scrE[(size_t)(it * R3 + k3) * T + tid], x[k3]);
            }
            bar_sync(BAR_MAIN, T);          // every warp is done reading the buffer

            // ---- half 1: O' = FFT_F(x'[2m+1]); join, x conj(T)/N, split into A (inverted now) and B (parked)
            pass1(std::true_type{}, std::true_type{}, i, 1, cispi(2.f * (turns0 + turns_h)), fs.rho, unused_energy);
            bar_sync(BAR_MAIN, T);
            if (THR_2X_SMEM_B && use_raw && tid == 0 && has_block(i + 2)) issue_tile(i + 2);      // for A(i+2)
            pass2();
            __syncwarp();
            // ---- per template: join, x conj(T)/N, the two inverse half-transforms, |c|^2 arg-max
#pragma unroll 1
            for (int tpl = 0; tpl < n_tpl; ++tpl) {
            const float2 *tlo = p.tpl_spec + (size_t)tpl * NB, *thi = tlo + F;
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int g = H::p3_item(tid, it);
                const uint32_t ab = (uint32_t)g * 136u;
                const int kb = (g >> LOG2R2) + 32 * (g & (R2 - 1));              // bin k = kb + S*k3 < F
                float2 wb = cispi(-2.0f * (float)kb / (float)NB);                // W_NB^kb
                // opaque: keeps the compiler from hoisting the 32 products wb * W_32^k3 out of the block loop
                // into a per-thread local-memory table (reloads would miss the small L1)
                asm volatile("" : "+f"(wb.x), "+f"(wb.y));
                float2 x[R3];
                if (!MULTI || tpl == 0) {    // pass 3 of O'; kept for the other templates
#pragma unroll
                    for (int n3 = 0; n3 < R3; ++n3) x[brev(n3, LOG2R3)] = ld8(ab + (uint32_t)n3 * 8u);
                    fft_dit<R3, false>(x);
                    if constexpr (MULTI) {
#pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3) __stcg(&scrO[(size_t)(it * R3 + k3) * T + tid], x[k3]);
                    }
                } else {
#pragma unroll
                    for (int k3 = 0; k3 < R3; ++k3) x[k3] = __ldcg(&scrO[(size_t)(it * R3 + k3) * T + tid]);
                }
                float2 y[R3];
                float2 ev[R3];           // all loads of E' first (no ld.cg moves above the st.cg of B below)
#pragma unroll
                for (int k3 = 0; k3 < R3; ++k3) ev[k3] = __ldcg(&scrE[(size_t)(it * R3 + k3) * T + tid]);
#pragma unroll
                for (int k3 = 0; k3 < R3; ++k3) {
                    const size_t sidx = (size_t)(it * R3 + k3) * T + tid;
                    // w = W_NB^(kb + S k3) = wb * W_32^k3
                    const float2 w = (k3 == 0) ? wb : mul_tw<false>(wb, cos32(k3), sin32(k3));
                    const float2 t = cmul(x[k3], w);
                    const float2 ylo = cmul(f2add(ev[k3], t), __ldg(&tlo[sidx]));
                    const float2 yhi = cmul(f2sub(ev[k3], t), __ldg(&thi[sidx]));
                    y[brev(k3, LOG2R3)] = f2add(ylo, yhi);
                    __stcg(&scrB[sidx], cmulc(f2sub(ylo, yhi), w));
                }
                fft_dit<R3, true>(y);
#pragma unroll
                for (int n3 = 0; n3 < R3; ++n3) st8(ab + (uint32_t)n3 * 8u, y[n3]);
            }

            // ---- two inverse half-transforms: par 0 -> even lags 2m, par 1 -> odd lags 2m+1
            uint32_t best_bits = 0u, best_lag = 0xffffffffu;
            float c1tot = 0.f, c2tot = 0.f;
#pragma unroll 1
            for (int e = 0; e < 2; ++e) {
                if (e == 1) {
                    // all pass-1' loads of the even half are done (arg-max barriers): restore B, inverse pass 3'
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const uint32_t ab = (uint32_t)H::p3_item(tid, it) * 136u;
                        float2 y[R3];
#pragma unroll
                        for (int k3 = 0; k3 < R3; ++k3)
                            y[brev(k3, LOG2R3)] = __ldcg(&scrB[(size_t)(it * R3 + k3) * T + tid]);
                        fft_dit<R3, true>(y);
#pragma unroll
                        for (int n3 = 0; n3 < R3; ++n3) st8(ab + (uint32_t)n3 * 8u, y[n3]);
                    }
                }
                __syncwarp();
                pass2_inv();
                bar_sync(BAR_MAIN, T);
                // inverse pass 1': conj twiddle on load, radix-32 over k1 -> c[2(n1*M + j) + e]
                float cp[32];
                float c1sum = 0.f, c2sum = 0.f, cbestv = 0.f;
                uint32_t inmask;
                {
                    float2 x[32];
                    float2 ws = w1, ws4 = w4;
                    asm volatile("" : "+f"(ws.x), "+f"(ws.y), "+f"(ws4.x), "+f"(ws4.y));
                    float2 cur[4];
                    cur[0] = ws;
                    cur[1] = cmul(ws, ws);
                    cur[2] = cmul(cur[1], ws);
                    cur[3] = ws4;
                    x[0] = ld8(a1b);
#pragma unroll
                    for (int k1 = 1; k1 < 32; ++k1) {
                        if (k1 > 4) cur[(k1 - 1) & 3] = cmul(cur[(k1 - 1) & 3], ws4);
                        x[brev(k1, 5)] = cmulc(ld8(a1b + (uint32_t)k1 * A1_STEP), cur[(k1 - 1) & 3]);
                    }
                    fft_dit<32, true>(x);
                    // lags 2m + e in [corr_start, corr_stop)  <=>  m in [start_e, stop_e)
                    const int start_e = (p.corr_start - e + 1) >> 1, stop_e = (p.corr_stop - e + 1) >> 1;
                    const int lo = max(0, (start_e - tid + M - 1) >> LOG2M);
                    const int hi = min(31, (stop_e - 1 - tid) >> LOG2M);
                    const uint32_t upto_hi = hi >= 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u);
                    inmask = (hi >= lo) ? (upto_hi & ~((1u << lo) - 1u)) : 0u;
                    float *cpe = cps + (size_t)e * F;
#pragma unroll
                    for (int n1 = 0; n1 < 32; ++n1) {
                        const float pv = x[n1].x * x[n1].x + x[n1].y * x[n1].y;
                        cp[n1] = pv;
                        if (inmask & (1u << n1)) cbestv = fmaxf(cbestv, pv);
                        __stcg(&cpe[n1 * M + tid], pv);
                    }
                    if (need_std_k) {
                        const int len_e = (p.corr_len - e + 1) >> 1;         // lags 2m + e < corr_len
#pragma unroll
                        for (int n1 = 0; n1 < 32; ++n1) {
                            if (n1 * M + tid < len_e) {
                                c1sum += sqrtf(cp[n1]);
                                c2sum += cp[n1];
                            }
                        }
                    }
                    if (p.dbg_corr && tpl == 0) {
#pragma unroll
                        for (int n1 = 0; n1 < 32; ++n1) {
                            const int lagd = 2 * (n1 * M + tid) + e;
                            if (lagd < p.corr_len) p.dbg_corr[lagd] = x[n1];
                        }
                    }
                }
                auto find_lag = [&](uint32_t gbits) {
                    uint32_t k = 0xffffffffu;
#pragma unroll
                    for (int n1 = 31; n1 >= 0; --n1)
                        if ((inmask & (1u << n1)) && __float_as_uint(cp[n1]) == gbits)
                            k = min(k, (uint32_t)(2 * (n1 * M + tid) + e));
                    return k;
                };
                ArgOut rb;
                if (need_std_k) rb = main_argmax<T, true>(__float_as_uint(cbestv), c1sum, c2sum, red, tid, find_lag);
                else rb = main_argmax<T, false>(__float_as_uint(cbestv), 0.f, 0.f, red, tid, find_lag);
                c1tot += rb.s0;
                c2tot += rb.s1;
                if (rb.vbits > best_bits || (rb.vbits == best_bits && rb.key < best_lag)) {
                    best_bits = rb.vbits;
                    best_lag = rb.key;
                }
            }
            if (tid == 0) {
                TailSlot &ts = tailslot[q * TPL_SLOTS + tpl];
                const int s = (int)best_lag;
                ts.peak_cp = __uint_as_float(best_bits);
                ts.s = s;
                ts.c1 = c1tot;
                ts.c2 = c2tot;
                // neighbours of the peak belong to the other half-transform: fetch |c|^2 from the scratch
                // (written before the arg-max barriers above, so visible to this thread)
                const int sm1 = s > 0 ? s - 1 : 0, sp1 = s + 1 < NB ? s + 1 : NB - 1;
                ts.pa = __ldcg(&cps[(size_t)(sm1 & 1) * F + (sm1 >> 1)]);
                ts.pc = __ldcg(&cps[(size_t)(sp1 & 1) * F + (sp1 >> 1)]);
            }
            }   // templates (the next one reuses the FFT buffer: all pass-1' loads are done, arg-max barriers)
            bar_arrive(BAR_TAILREQ + q, NTHREADS);
        }
    }
#undef use_raw
#undef need_std_k
}

}  // namespace thr
