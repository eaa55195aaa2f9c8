// identify.cu -- the step after `detect` on the GPU: transmitter ids from the carrier frequency and the mask that drops
// duplicate detections of one transmission in adjacent blocks (thrifty/identify.py:26-166).
//
// The reference works on Python lists; at 10^7 detections a second that is the next bottleneck behind the detect kernel.
// Here the detections are columns (one value per detection, any order, any number of receivers):
//   * classification by frequency map (identify.py:106-118): one thread per detection, last matching range wins;
//   * automatic classification (identify.py:26-103): carrier-bin histogram per receiver with atomics; the scan of the
//     ~100 histogram bins for peaks is sequential by nature and stays with the caller; np.digitize on the device;
//   * duplicates (identify.py:134-164): np.argsort over (rxid, txid, block, timestamp) becomes a bitonic sort of 192-bit
//     keys (made unique by the detection's position, so no stability question arises), then every detection looks at its
//     neighbours in sorted order -- np.roll semantics, i.e. cyclic, and without checking that the neighbour belongs to
//     the same transmitter, exactly as the reference does.
// Host pointers in, host pointers out; the data is tiny (tens of bytes per detection), the copies are not the point.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/thrifty_b200.h"

namespace {

thread_local std::string g_identify_error;

struct SortKey {
    uint64_t k0;      // (rxid, txid) as order-preserving unsigned
    uint64_t k1;      // block
    uint64_t k2;      // timestamp, order-preserving bits of the double
    uint32_t idx;     // position of the detection in the input (0xffffffff: padding)
    uint32_t pad;
};

__device__ __forceinline__ bool key_less(const SortKey &a, const SortKey &b) {
    if (a.k0 != b.k0) return a.k0 < b.k0;
    if (a.k1 != b.k1) return a.k1 < b.k1;
    if (a.k2 != b.k2) return a.k2 < b.k2;
    return a.idx < b.idx;
}

__global__ void classify_kernel(int64_t n, const int32_t *rxid, const int32_t *cbin, const double *coff, int n_map,
                                const int32_t *m_rx, const int32_t *m_tx, const double *m_lo, const double *m_hi,
                                int32_t *txid) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double f = (double)cbin[i] + coff[i];          // identify.py:110
    int32_t t = -1;
    for (int m = 0; m < n_map; ++m)
        if (m_rx[m] == rxid[i] && f >= m_lo[m] && f <= m_hi[m]) t = m_tx[m];   // the last matching range wins (:113-116)
    txid[i] = t;
}

__global__ void histogram_kernel(int64_t n, const int32_t *rxid, const int32_t *cbin, int32_t which_rx, int32_t first_bin,
                                 int32_t n_bins, unsigned int *counts) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || rxid[i] != which_rx) return;
    const int32_t b = cbin[i] - first_bin;
    if (b >= 0 && b < n_bins) atomicAdd(&counts[b], 1u);
}

__global__ void minmax_kernel(int64_t n, const int32_t *rxid, const int32_t *cbin, int32_t which_rx, int32_t *mn, int32_t *mx) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || rxid[i] != which_rx) return;
    atomicMin(mn, cbin[i]);
    atomicMax(mx, cbin[i]);
}

// np.digitize(bin, edges) - 1 with increasing edges (identify.py:98-99): number of edges <= bin, minus one
__global__ void digitize_kernel(int64_t n, const int32_t *rxid, const int32_t *cbin, int32_t which_rx, int n_edges,
                                const int64_t *edges, int32_t *txid) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || rxid[i] != which_rx) return;
    int c = 0;
    for (int e = 0; e < n_edges; ++e) c += (edges[e] <= (int64_t)cbin[i]) ? 1 : 0;
    txid[i] = c - 1;
}

__global__ void build_keys_kernel(int64_t n, int64_t n_pad, const int32_t *rxid, const int32_t *txid, const int32_t *block,
                                  const double *ts, SortKey *keys) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    SortKey k;
    if (i < n) {
        k.k0 = ((uint64_t)((uint32_t)rxid[i] ^ 0x80000000u) << 32) | (uint64_t)((uint32_t)txid[i] ^ 0x80000000u);
        k.k1 = (uint64_t)((uint32_t)block[i] ^ 0x80000000u);
        const uint64_t b = (uint64_t)__double_as_longlong(ts[i]);
        k.k2 = (b >> 63) ? ~b : (b | 0x8000000000000000ull);     // total order of doubles as unsigned
        k.idx = (uint32_t)i;
    } else {
        k.k0 = k.k1 = k.k2 = ~0ull;                               // padding sorts last
        k.idx = 0xffffffffu;
    }
    k.pad = 0;
    keys[i] = k;
}

__global__ void bitonic_step_kernel(SortKey *keys, int64_t n_pad, int64_t j, int64_t k) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const int64_t l = i ^ j;
    if (l <= i) return;
    const SortKey a = keys[i], b = keys[l];
    const bool up = (i & k) == 0;
    if (up ? key_less(b, a) : key_less(a, b)) {
        keys[i] = b;
        keys[l] = a;
    }
}

// identify.py:151-160 on the sorted order: cur / np.roll(cur, 1) / np.roll(cur, -1)
__global__ void duplicate_mask_kernel(int64_t n, const SortKey *keys, const int32_t *txid, const int32_t *block,
                                      const double *energy, uint8_t *keep) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = keys[i].idx, p = keys[(i + n - 1) % n].idx, q = keys[(i + 1) % n].idx;
    const bool unidentified = txid[c] == -1;
    const bool dup_prev = (block[c] == block[p] + 1) && (energy[c] < energy[p]);
    const bool dup_next = (block[c] == block[q] - 1) && (energy[c] < energy[q]);
    keep[c] = (dup_prev || dup_next || unidentified) ? 0 : 1;
}

int id_fail(const char *what, cudaError_t e) {
    g_identify_error = std::string(what) + ": " + cudaGetErrorString(e);
    return THR_ERR_CUDA;
}

#define IDCU(call)                                              \
    do {                                                        \
        cudaError_t e_ = (call);                                \
        if (e_ != cudaSuccess) { free_all(); return id_fail(#call, e_); } \
    } while (0)

struct DevBufs {
    std::vector<void *> ptrs;
    template <class T>
    cudaError_t up(T **dst, const T *src, size_t n) {
        cudaError_t e = cudaMalloc((void **)dst, (n ? n : 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(*dst);
        return src ? cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
    }
    void free_all() {
        for (void *p : ptrs) cudaFree(p);
        ptrs.clear();
    }
};

inline unsigned grid_for(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" {

const char *thr_identify_last_error(void) { return g_identify_error.c_str(); }

int thr_identify_classify(int32_t device, int64_t n, const int32_t *rxid, const int32_t *carrier_bin,
                          const double *carrier_offset, int32_t n_map, const int32_t *map_rxid, const int32_t *map_txid,
                          const double *map_start, const double *map_stop, int32_t *txid_out) {
    if (n < 0 || n_map < 0 || (n > 0 && (!rxid || !carrier_bin || !carrier_offset || !txid_out))) return THR_ERR_INVALID;
    if (n == 0) return THR_OK;
    DevBufs B;
    auto free_all = [&] { B.free_all(); };
    IDCU(cudaSetDevice(device));
    int32_t *d_rx, *d_bin, *d_tx, *d_mrx, *d_mtx;
    double *d_off, *d_lo, *d_hi;
    IDCU(B.up(&d_rx, rxid, (size_t)n));
    IDCU(B.up(&d_bin, carrier_bin, (size_t)n));
    IDCU(B.up(&d_off, carrier_offset, (size_t)n));
    IDCU(B.up(&d_mrx, map_rxid, (size_t)n_map));
    IDCU(B.up(&d_mtx, map_txid, (size_t)n_map));
    IDCU(B.up(&d_lo, map_start, (size_t)n_map));
    IDCU(B.up(&d_hi, map_stop, (size_t)n_map));
    IDCU(B.up(&d_tx, (const int32_t *)nullptr, (size_t)n));
    classify_kernel<<<grid_for(n), 256>>>(n, d_rx, d_bin, d_off, n_map, d_mrx, d_mtx, d_lo, d_hi, d_tx);
    IDCU(cudaGetLastError());
    IDCU(cudaMemcpy(txid_out, d_tx, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    free_all();
    return THR_OK;
}

int thr_identify_bin_histogram(int32_t device, int64_t n, const int32_t *rxid, const int32_t *carrier_bin, int32_t which_rxid,
                               int32_t *first_bin, int32_t *n_bins, uint32_t *counts, int32_t counts_cap) {
    if (n < 0 || !first_bin || !n_bins || (n > 0 && (!rxid || !carrier_bin))) return THR_ERR_INVALID;
    *n_bins = 0;
    if (n == 0) return THR_OK;
    DevBufs B;
    auto free_all = [&] { B.free_all(); };
    IDCU(cudaSetDevice(device));
    int32_t *d_rx, *d_bin, *d_mm;
    unsigned int *d_cnt;
    IDCU(B.up(&d_rx, rxid, (size_t)n));
    IDCU(B.up(&d_bin, carrier_bin, (size_t)n));
    const int32_t init[2] = {INT32_MAX, INT32_MIN};
    IDCU(B.up(&d_mm, init, 2));
    minmax_kernel<<<grid_for(n), 256>>>(n, d_rx, d_bin, which_rxid, d_mm, d_mm + 1);
    int32_t mm[2];
    IDCU(cudaMemcpy(mm, d_mm, sizeof mm, cudaMemcpyDeviceToHost));
    if (mm[0] > mm[1]) { free_all(); return THR_OK; }              // this receiver has no detections
    const int64_t nb = (int64_t)mm[1] - mm[0] + 1;
    *first_bin = mm[0];
    *n_bins = (int32_t)nb;
    if (!counts || nb > counts_cap) { free_all(); return counts ? THR_ERR_NOMEM : THR_OK; }   // size query
    IDCU(B.up(&d_cnt, (const unsigned int *)nullptr, (size_t)nb));
    IDCU(cudaMemset(d_cnt, 0, (size_t)nb * sizeof(unsigned int)));
    histogram_kernel<<<grid_for(n), 256>>>(n, d_rx, d_bin, which_rxid, mm[0], (int32_t)nb, d_cnt);
    IDCU(cudaGetLastError());
    IDCU(cudaMemcpy(counts, d_cnt, (size_t)nb * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    free_all();
    return THR_OK;
}

int thr_identify_digitize(int32_t device, int64_t n, const int32_t *rxid, const int32_t *carrier_bin, int32_t which_rxid,
                          int32_t n_edges, const int64_t *edges, int32_t *txid_inout) {
    if (n < 0 || n_edges < 0 || (n > 0 && (!rxid || !carrier_bin || !txid_inout)) || (n_edges > 0 && !edges)) return THR_ERR_INVALID;
    if (n == 0) return THR_OK;
    DevBufs B;
    auto free_all = [&] { B.free_all(); };
    IDCU(cudaSetDevice(device));
    int32_t *d_rx, *d_bin, *d_tx;
    int64_t *d_e;
    IDCU(B.up(&d_rx, rxid, (size_t)n));
    IDCU(B.up(&d_bin, carrier_bin, (size_t)n));
    IDCU(B.up(&d_tx, txid_inout, (size_t)n));
    IDCU(B.up(&d_e, edges, (size_t)n_edges));
    digitize_kernel<<<grid_for(n), 256>>>(n, d_rx, d_bin, which_rxid, n_edges, d_e, d_tx);
    IDCU(cudaGetLastError());
    IDCU(cudaMemcpy(txid_inout, d_tx, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    free_all();
    return THR_OK;
}

int thr_identify_duplicates(int32_t device, int64_t n, const int32_t *rxid, const int32_t *txid, const int32_t *block,
                            const double *timestamp, const double *energy, uint8_t *keep_out) {
    if (n < 0 || n >= 0x7fffffff || (n > 0 && (!rxid || !txid || !block || !timestamp || !energy || !keep_out))) return THR_ERR_INVALID;
    if (n == 0) return THR_OK;
    DevBufs B;
    auto free_all = [&] { B.free_all(); };
    IDCU(cudaSetDevice(device));
    int64_t n_pad = 1;
    while (n_pad < n) n_pad <<= 1;
    int32_t *d_rx, *d_tx, *d_blk;
    double *d_ts, *d_en;
    SortKey *d_keys;
    uint8_t *d_keep;
    IDCU(B.up(&d_rx, rxid, (size_t)n));
    IDCU(B.up(&d_tx, txid, (size_t)n));
    IDCU(B.up(&d_blk, block, (size_t)n));
    IDCU(B.up(&d_ts, timestamp, (size_t)n));
    IDCU(B.up(&d_en, energy, (size_t)n));
    IDCU(B.up(&d_keys, (const SortKey *)nullptr, (size_t)n_pad));
    IDCU(B.up(&d_keep, (const uint8_t *)nullptr, (size_t)n));
    build_keys_kernel<<<grid_for(n_pad), 256>>>(n, n_pad, d_rx, d_tx, d_blk, d_ts, d_keys);
    for (int64_t k = 2; k <= n_pad; k <<= 1)
        for (int64_t j = k >> 1; j > 0; j >>= 1)
            bitonic_step_kernel<<<grid_for(n_pad), 256>>>(d_keys, n_pad, j, k);
    duplicate_mask_kernel<<<grid_for(n), 256>>>(n, d_keys, d_tx, d_blk, d_en, d_keep);
    IDCU(cudaGetLastError());
    IDCU(cudaMemcpy(keep_out, d_keep, (size_t)n, cudaMemcpyDeviceToHost));
    free_all();
    return THR_OK;
}

}  // extern "C"
