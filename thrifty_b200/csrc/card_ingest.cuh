// card_ingest.cuh -- `.card` text ingest: line scan on the host, base64 decode on the GPU.
//
// A `.card` data line is "<sec>.<usec> <block_idx> <base64 of 2N I/Q bytes>\n"
// (fastcard/fastcard_cli.c:184-192); readers: thrifty/block_data.py:101-131 (Python, lenient)
// and fastcard/card_reader.c:22-78 (C, strict about the payload length).  The host only finds
// line boundaries and parses the two numbers (~30 bytes of every 43.7 KB line); the payload
// is copied to the device as text and decoded there (43.7 KB read + 32.8 KB written per block:
// HBM-bound byte work, negligible next to the detect kernel), so the host never touches base64.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace thr {

// 6-bit value of a base64 character (standard alphabet); 64 for '=', 0xff for anything else
__device__ __forceinline__ uint32_t b64_value(uint32_t c) {
    uint32_t v = 0xffu;
    v = (c - 'A' < 26u) ? c - 'A' : v;
    v = (c - 'a' < 26u) ? c - 'a' + 26u : v;
    v = (c - '0' < 10u) ? c - '0' + 52u : v;
    v = (c == '+') ? 62u : v;
    v = (c == '/') ? 63u : v;
    v = (c == '=') ? 64u : v;
    return v;
}

// One CTA per (block, segment of SEG_CHARS characters).  The payload starts at an arbitrary byte
// offset: it is fetched with aligned 32-bit loads and re-aligned while being staged in shared
// memory; each thread then decodes 16 characters into 12 bytes.
constexpr int B64_SEG_CHARS = 4096;          // characters per CTA  -> 3072 output bytes
constexpr int B64_THREADS = 256;             // 16 characters per thread

__global__ void __launch_bounds__(B64_THREADS)
b64_decode_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ payload_off, int n_chars,
                  int raw_bytes, uint8_t *__restrict__ raw_out, unsigned int *__restrict__ n_bad, int blk_base) {
    __shared__ __align__(16) uint8_t chars[B64_SEG_CHARS + 16];
    const int blk = blockIdx.y;
    const int seg0 = blockIdx.x * B64_SEG_CHARS;
    const int seg_len = min(B64_SEG_CHARS, n_chars - seg0);
    const uint8_t *src = text + payload_off[blk] + seg0;
    const uintptr_t mis = reinterpret_cast<uintptr_t>(src) & 3u;
    const uint32_t *src32 = reinterpret_cast<const uint32_t *>(src - mis);
    const int n_words = (int)((seg_len + mis + 3) >> 2);
    for (int w = threadIdx.x; w < n_words; w += B64_THREADS) {
        const uint32_t v = __ldg(&src32[w]);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int pos = 4 * w + b - (int)mis;
            if (pos >= 0 && pos < seg_len) chars[pos] = (uint8_t)(v >> (8 * b));
        }
    }
    __syncthreads();
    const int c0 = threadIdx.x * 16;
    if (c0 >= seg_len) return;
    uint32_t out[3];
    uint32_t bad = 0;
    int valid_bytes = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {                 // 4 groups of 4 characters -> 3 bytes each
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int pos = c0 + 4 * g + k;
            v[k] = pos < seg_len ? b64_value(chars[pos]) : 64u;
        }
        const int npad = (v[3] == 64u) + (v[3] == 64u && v[2] == 64u);
        if (c0 + 4 * g < seg_len) {
            bad |= (v[0] >= 64u) | (v[1] >= 64u) | (v[2] > 64u) | (v[3] > 64u);
            bad |= (npad != 0) & (seg0 + c0 + 4 * g + 4 < n_chars);      // '=' only in the last group
            valid_bytes += 3 - npad;
        }
        const uint32_t w24 = ((v[0] & 63u) << 18) | ((v[1] & 63u) << 12) | ((v[2] & 63u) << 6) | (v[3] & 63u);
        const uint32_t b0 = (w24 >> 16) & 255u, b1 = (w24 >> 8) & 255u, b2 = w24 & 255u;
        // pack the 12 output bytes little-endian into three 32-bit words
        if (g == 0) { out[0] = b0 | (b1 << 8) | (b2 << 16); }
        if (g == 1) { out[0] |= b0 << 24; out[1] = b1 | (b2 << 8); }
        if (g == 2) { out[1] |= (b0 << 16) | (b1 << 24); out[2] = b2; }
        if (g == 3) { out[2] |= (b0 << 8) | (b1 << 16) | (b2 << 24); }
    }
    if (bad) {      // n_bad[0] = count; the first reporter also leaves (block, character position, 16 raw characters)
        if (atomicAdd(n_bad, 1u) == 0u) {
            n_bad[1] = (unsigned int)(blk_base + blk);
            n_bad[2] = (unsigned int)(seg0 + c0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned int w = 0;
                for (int b = 0; b < 4; ++b) w |= (unsigned int)(c0 + 4 * k + b < seg_len ? chars[c0 + 4 * k + b] : 0) << (8 * b);
                n_bad[3 + k] = w;
            }
        }
    }
    const int o0 = (seg0 / 4) * 3 + threadIdx.x * 12;          // byte offset inside the block's raw row
    uint8_t *dst = raw_out + (size_t)blk * raw_bytes + o0;
    if (valid_bytes == 12 && o0 + 12 <= raw_bytes) {
        uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);     // rows and o0 are multiples of 4
        d32[0] = out[0];
        d32[1] = out[1];
        d32[2] = out[2];
    } else {
        for (int b = 0; b < valid_bytes && o0 + b < raw_bytes; ++b) dst[b] = (uint8_t)(out[b >> 2] >> (8 * (b & 3)));
    }
}

}  // namespace thr
