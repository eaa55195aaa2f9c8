// overlap.cu -- does the FMA pipe (packed FP32x2) overlap with the shared-memory pipe on sm_100?
//
// Mimics one pass of the detect kernel: per iteration every thread loads 32 complex values (LDS.64) from a padded
// shared buffer, runs NFMA packed FMAs on them (32 independent chains) and stores 32 values back (STS.64).
//   mode 0: compute only            mode 1: load/store only
//   mode 2: both, CTA barrier every iteration (all 16 warps phase-locked, as between FFT passes)
//   mode 3: both, no barrier (warps drift)      mode 4: both, no barrier, odd warps start half an iteration late
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o overlap overlap.cu && ./overlap
#include <cstdio>
#include <cuda_runtime.h>

constexpr int T = 512, ITERS = 200;

template <int MODE, int NFMA>
__global__ void __launch_bounds__(T, 1) k(float2 *out, float seed) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    // padded rows of 16 complex (136 B): thread reads column (tid & 15) of rows (tid >> 4) + 32*i  (pass-2 pattern)
    unsigned char *base = smem + (tid >> 4) * 136 + (tid & 15) * 8;
    unsigned char *base2 = smem + (tid >> 4) * 136 + ((tid + 1) & 15) * 8;      // neighbouring column: a real copy
    for (int i = tid; i < 139264 / 8; i += T) reinterpret_cast<float2 *>(smem)[i] = make_float2(seed * i, 1.0f);
    __syncthreads();
    float2 x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = make_float2(seed + i, seed - i);
    const float2 w = make_float2(0.999f + seed, 0.998f), c = make_float2(seed, 2.f * seed);
    auto compute = [&]() {
#pragma unroll
        for (int r = 0; r < NFMA / 32; ++r)
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = __ffma2_rn(x[i], w, c);
    };
    if (MODE == 4 && ((tid >> 5) & 1)) compute();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float2 v = *reinterpret_cast<const float2 *>(base + i * 4352);
                x[i] = MODE == 1 ? v : __fadd2_rn(x[i], v);
            }
        }
        if (MODE != 1) compute();
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) *reinterpret_cast<float2 *>(base2 + i * 4352) = x[i];
        }
        if (MODE == 2 || MODE == 1) __syncthreads();      // mode 1 needs it for correctness of the copy chain
    }
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 32; ++i) { acc.x += x[i].x; acc.y += x[i].y; }
    out[blockIdx.x * T + tid] = acc;
}

// Same work per SM on 32 warps: 1024 threads x 16 values, the radix-32 item split over a thread pair that exchanges
// 8 complex values (16 SHFL) per pass; 64 registers per thread.
template <int MODE, int NFMA>
__global__ void __launch_bounds__(1024, 1) k1024(float2 *out, float seed) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int item = tid >> 1, half = tid & 1;
    unsigned char *base = smem + (item >> 4) * 136 + (item & 15) * 8 + half * 16 * 4352;
    unsigned char *base2 = smem + (item >> 4) * 136 + ((item + 1) & 15) * 8 + half * 16 * 4352;
    for (int i = tid; i < 139264 / 8; i += 1024) reinterpret_cast<float2 *>(smem)[i] = make_float2(seed * i, 1.0f);
    __syncthreads();
    float2 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = make_float2(seed + i, seed - i);
    const float2 w = make_float2(0.999f + seed, 0.998f), c = make_float2(seed, 2.f * seed);
    for (int it = 0; it < ITERS; ++it) {
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fadd2_rn(x[i], *reinterpret_cast<const float2 *>(base + i * 4352));
        }
#pragma unroll
        for (int r = 0; r < NFMA / 32; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __ffma2_rn(x[i], w, c);
            if (r == NFMA / 64) {           // pair exchange in the middle of the item: 8 complex = 16 shuffles
#pragma unroll
                for (int i = 0; i < 8; ++i) {       // selects, not a dynamic register index
                    const float2 send = half ? x[i] : x[8 + i];
                    float2 recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
                    x[i] = half ? recv : x[i];
                    x[8 + i] = half ? x[8 + i] : recv;
                }
            }
        }
        if (MODE != 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) *reinterpret_cast<float2 *>(base2 + i * 4352) = x[i];
        }
        if (MODE == 2) __syncthreads();
    }
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc.x += x[i].x; acc.y += x[i].y; }
    out[blockIdx.x * 1024 + tid] = acc;
}

template <int MODE, int NFMA>
float run1024(float2 *d_out) {
    auto fn = k1024<MODE, NFMA>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 139264);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fn<<<148, 1024, 139264>>>(d_out, 1e-6f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    fn<<<148, 1024, 139264>>>(d_out, 1e-6f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

template <int MODE, int NFMA>
float run(float2 *d_out) {
    auto fn = k<MODE, NFMA>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 139264);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fn<<<148, T, 139264>>>(d_out, 1e-6f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    fn<<<148, T, 139264>>>(d_out, 1e-6f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

template <int NFMA>
void suite(float2 *d_out, double ghz) {
    const float t0 = run<0, NFMA>(d_out), t1 = run<1, NFMA>(d_out), t2 = run<2, NFMA>(d_out), t3 = run<3, NFMA>(d_out),
                t4 = run<4, NFMA>(d_out);
    auto cyc = [&](float ms) { return ms * 1e-3 * ghz * 1e9 / ITERS; };
    printf("{\"packed_fma_per_thread_per_iter\": %d, \"cycles_per_iter\": {\"compute_only\": %.0f, \"lds_sts_only\": %.0f, "
           "\"both_barrier\": %.0f, \"both_free\": %.0f, \"both_skewed\": %.0f}, \"sum\": %.0f, \"max\": %.0f}\n",
           NFMA, cyc(t0), cyc(t1), cyc(t2), cyc(t3), cyc(t4), cyc(t0) + cyc(t1), cyc(t0) > cyc(t1) ? cyc(t0) : cyc(t1));
    const float u0 = run1024<0, NFMA>(d_out), u2 = run1024<2, NFMA>(d_out), u3 = run1024<3, NFMA>(d_out);
    printf("{\"packed_fma_per_item_per_iter\": %d, \"threads\": 1024, \"cycles_per_iter\": {\"compute_and_shuffles_only\": %.0f, "
           "\"both_barrier\": %.0f, \"both_free\": %.0f}}\n", NFMA, cyc(u0), cyc(u2), cyc(u3));
}

int main() {
    float2 *d_out;
    cudaMalloc(&d_out, 148 * 1024 * sizeof(float2));
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("{\"sm_clock_ghz\": %.3f, \"threads\": %d, \"note\": \"32 LDS.64 + N FFMA2 + 32 STS.64 per thread per iteration, 16 warps/SM\"}\n", ghz, T);
    suite<128>(d_out, ghz);
    suite<256>(d_out, ghz);
    suite<352>(d_out, ghz);
    suite<512>(d_out, ghz);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
