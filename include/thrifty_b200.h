/*
 * thrifty_b200.h -- C ABI of the B200-native Thrifty `detect` hot path.
 *
 * One shared object (libthrifty_b200.so, built by nvcc for sm_100a) exposes the
 * per-block detect chain of swkrueger/Thrifty:
 *
 *   uint8 IQ -> complex64 -> FFT -> windowed spectral-peak carrier detect ->
 *   Dirichlet sub-bin fit -> time-domain mix -> FFT -> x conj(FFT(template)) ->
 *   IFFT -> |corr| windowed peak + threshold -> Gaussian sub-sample offset -> SoA
 *
 * fused into a single persistent CUDA kernel (one launch per batch of blocks).
 * No torch / numpy / cuFFT types cross this boundary: plain pointers and sizes.
 *
 * Reference interfaces this ABI replaces (paths inside the reference checkout):
 *   - thrifty/detect.py:34-91        class Detector (ctor + detect())         -> thr_create / thr_detect_batch*
 *   - thrifty/carrier_sync.py:82-118 DefaultSynchronizer                      -> fused (carrier stage of the kernel)
 *   - thrifty/soa_estimator.py:42-124 SoaEstimator                            -> fused (correlation stage of the kernel)
 *   - fastdet/corr_detector.h:24-45  CorrDetector(template, block_len, history_len, thresh...) / detect()
 *   - fastcard/fastcard.h:46-53      fastcard_new / fastcard_process / fastcard_free (handle + int status model)
 *
 * Conventions (mirroring fastcard.h): opaque handle, int status (0 ok, <0 error),
 * no exceptions, caller owns input/output buffers, the handle owns device scratch
 * and the template spectra, one handle = one host thread + one CUDA stream.
 */
#ifndef THRIFTY_B200_H
#define THRIFTY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THR_ABI_VERSION 1

/* status codes */
#define THR_OK             0
#define THR_ERR_INVALID   -1   /* bad argument / unsupported configuration   */
#define THR_ERR_CUDA      -2   /* CUDA runtime error (see thr_last_error)    */
#define THR_ERR_NOMEM     -3
#define THR_ERR_NO_DEVICE -4   /* no usable sm_100 device                    */

/* thr_config.flags */
#define THR_CFG_OVERLAP_LAUNCHES 1u    /* consecutive *_device launches on one stream may overlap at their
                                          edges (programmatic dependent launch): the next batch starts on SMs
                                          the previous batch has drained.  Launches are independent, so this is
                                          safe as long as consecutive launches do not write the same output
                                          buffer (alternate two record buffers).                              */

#define THR_CFG_FASTDET_SEMANTICS 2u    /* semantics of the reference's native twin instead of the Python path:
                                          fastcard/cardet.c:7-41 + fastdet/corr_detector.cpp:88-197.  Decisions
                                          on powers (threshold = c + s*noise_power, stddev terms must be 0),
                                          integer-bin carrier shift (the spectrum roll is folded into the
                                          template: 2 transforms per block instead of 3), 3-point parabolic
                                          carrier offset (reporting only), Gaussian correlation offset clipped
                                          to +-0.5, noise clamped at 0, no zero-straddling carrier window, one
                                          template.  Record fields keep their meaning (magnitudes = sqrt of the
                                          powers, as fastdet prints them, fastdet.cpp:191-206).               */

#define THR_CFG_GENERIC_KERNEL 4u       /* block_len 32768: always run the generic global-scratch kernel instead of
                                          the 2 x 16384 shared-memory kernel (tests / comparisons)              */

/* thr_record.flags */
#define THR_FLAG_CARRIER_DETECTED 1u   /* carrier peak above threshold (carrier_sync.py:69)      */
#define THR_FLAG_CORR_DETECTED    2u   /* correlation peak above threshold (soa_estimator.py:85) */

typedef struct thr_detector thr_detector;

/*
 * Detector configuration == thrifty/detect.py:24-31 DetectorSettings plus
 * launch geometry.  `templates` is borrowed only for the duration of thr_create.
 */
typedef struct thr_config {
    int32_t block_len;        /* N, power of two, 1024..32768                            */
    int32_t history_len;      /* H, samples repeated from the previous block             */
    int32_t template_len;     /* L, samples per template; L <= N and H >= L-1            */
    int32_t n_templates;      /* T >= 1 templates, stored back to back [T][L]            */
    const double *templates;  /* real template samples (numpy .npy float64 layout)       */
    int32_t carrier_len;      /* W of the Dirichlet kernel (detect.py:207: len(template))*/
    int32_t window_start;     /* carrier window, signed bins, CLOSED interval            */
    int32_t window_stop;      /*   (carrier_detect.py:17-58; 0,-1 = whole spectrum)      */
    double carrier_thresh[3]; /* (constant, snr, stddev) -- setting_parsers.py:141-185   */
    double corr_thresh[3];    /* (constant, snr, stddev)                                 */
    int32_t device;           /* CUDA device ordinal                                     */
    int32_t max_batch;        /* max blocks per launch (device staging is sized for it)  */
    uint32_t flags;           /* THR_CFG_* bits                                          */
    int32_t reserved;
} thr_config;

/*
 * One record per (block, template): the fields of thrifty/toads_data.py:8-45
 * (DetectionResult / CarrierSyncInfo / CorrDetectionInfo) as a 64-byte POD.
 * timestamp / rxid stay on the host.  If the carrier is not detected the corr_*
 * fields are: corr_sample = -1, others NaN, soa = NaN.
 */
typedef struct thr_record {
    int64_t block_idx;        /* echo of the input block index                           */
    double  soa;              /* (N-H)*block_idx + corr_sample + corr_offset (detect.py:67) */
    int32_t carrier_bin;      /* FFT index of the carrier peak                           */
    float   carrier_offset;   /* Dirichlet-fit sub-bin offset (0 if no carrier)          */
    float   carrier_energy;   /* |X[bin]| (magnitude, as in the .toad column)            */
    float   carrier_noise;    /* carrier noise rms estimate                              */
    int32_t corr_sample;      /* correlation peak lag                                    */
    float   corr_offset;      /* Gaussian sub-sample offset, clipped to +-0.6            */
    float   corr_energy;      /* |corr[sample]|                                          */
    float   corr_noise;       /* correlation noise rms estimate (may be NaN)             */
    uint32_t flags;           /* THR_FLAG_*                                              */
    int32_t template_idx;     /* which template this record belongs to                   */
    float   signal_energy;    /* mean |X'|^2 == sum |x|^2 (soa_estimator.py:111)         */
    float   reserved;
} thr_record;

typedef struct thr_info {
    int32_t abi_version;
    int32_t device;
    int32_t sm_count;
    int32_t grid;             /* CTAs per launch (persistent)                            */
    int32_t threads;          /* threads per CTA                                         */
    int32_t smem_bytes;       /* dynamic shared memory per CTA                           */
    int32_t ctas_per_sm;
    int32_t buffer_in_smem;   /* 1: FFT working set in shared memory; 0: global scratch  */
    int64_t launches;         /* detect-kernel launches issued by this handle so far     */
    char    device_name[64];
    char    kernel[64];
} thr_info;

/* ---- lifecycle (cf. fastcard_new / fastcard_free, fastcard/fastcard.h:46-53; CorrDetector's constructor,
 * fastdet/corr_detector.h:24-45; thrifty/detect.py:40-58 Detector.__init__) ----
 * thr_create fails with THR_ERR_NO_DEVICE unless the device is sm_100 (the library carries sm_100a code only; there is
 * no CPU fallback).  Testing aid: the environment variable THRIFTY_B200_MAX_GRID=<n> caps the persistent grid at n CTAs
 * so that small inputs exercise the multi-block pipeline of a CTA (compute-sanitizer runs).
 * THRIFTY_B200_COPY_THREADS=<1..16> sets how many host threads copy pageable input into the page-locked staging buffers
 * (default 8, or 4 on hosts with fewer than 16 hardware threads); they use non-temporal stores unless
 * THRIFTY_B200_COPY_NT=0. */
int  thr_create(const thr_config *cfg, thr_detector **out);
void thr_destroy(thr_detector *det);
/* Message for the last error on `det`, or for the last failed thr_create if det == NULL. */
const char *thr_last_error(const thr_detector *det);
int  thr_get_info(const thr_detector *det, thr_info *info);
int  thr_device_count(void);

/* ---- detection, host buffers (copies included; the reference-facing call) ----
 * raw:       n_blocks * 2N uint8, interleaved I,Q (the payload of a .card line,
 *            block_data.py:129-131 / fastcard/card_reader.c:69-75)
 * block_idx: n_blocks int64 block indices (NULL -> 0,1,2,...)
 * out:       n_blocks * n_templates records
 * n_blocks may exceed max_batch; the call chunks and overlaps copies with compute.
 * Host buffers may be pageable or page-locked (thr_host_alloc): page-locked ones are DMA'd directly (PCIe-bound,
 * ~1.6 M blocks/s at N = 16384), pageable ones go through internal page-locked staging (host-memory bound, ~0.7-1.2 M blocks/s);
 * records always return through page-locked staging.  The same holds for thr_detect_card and thr_detect_stream. */
int thr_detect_batch(thr_detector *det, const uint8_t *raw, const int64_t *block_idx,
                     int64_t n_blocks, thr_record *out);
/* Same, complex64 samples (interleaved re,im float32), n_blocks * N * 8 bytes:
 * the type the Python seam passes (detect.py:60-62). */
int thr_detect_batch_c64(thr_detector *det, const float *iq, const int64_t *block_idx,
                         int64_t n_blocks, thr_record *out);

/* ---- `.card` text ingest (thrifty/block_data.py:101-131 card_reader, fastcard/card_reader.c:22-78) ----
 * thr_card_scan: host-side line scan of a chunk of `.card` text.  Comment ('#'), blank and the
 *   'Using Volk machine:' / 'linux;' noise lines are skipped; for each data line
 *   "<time> <block_idx> <base64>" the two numbers are parsed and the byte offset of the payload is
 *   returned.  A payload whose length is not 4*ceil(2N/3) is an error (*bad_line = 1-based line number).
 *   With final_chunk == 0 an unterminated last line is left unconsumed (*consumed = bytes used).
 * thr_detect_card: scan + copy the text to the device + base64 decode ON THE GPU + detect, pipelined
 *   in chunks (the lines of chunk c+1 are scanned while chunk c crosses PCIe); timestamps / block_idx / out must hold
 *   max_blocks entries (out: max_blocks * n_templates).  On a malformed line the call fails with the 1-based line
 *   number in thr_last_error; *n_blocks / *consumed then tell how far the scan got and `out` is unspecified. */
int thr_card_scan(const char *text, size_t len, int32_t block_len, int32_t final_chunk, int64_t max_blocks,
                  double *timestamps, int64_t *block_idx, int64_t *payload_off, int64_t *n_found,
                  int64_t *consumed, int64_t *bad_line);
int thr_detect_card(thr_detector *det, const char *text, size_t len, int32_t final_chunk, int64_t max_blocks,
                    double *timestamps, int64_t *block_idx, thr_record *out, int64_t *n_blocks,
                    int64_t *consumed);

/* ---- contiguous raw sample streams (thrifty/block_data.py:70-98 block_reader, fastcard/raw_reader.c:15-46) ----
 * Block b covers samples [b*(N-H) - H, b*(N-H) + N - H) of the stream; the kernel reads these overlapping
 * windows in place, so only the N-H new samples of a block are ever copied.  `stream` starts with the H
 * history samples of block `first_block`; the number of whole blocks found is returned in *n_blocks.
 * Requires 2*(N-H) to be a multiple of 16 bytes (TMA bulk copy alignment).
 * The very first block of a capture (history = zeros, not representable as bytes) is handled by the caller
 * through thr_detect_batch_c64 (thrifty_b200.block_data.block_reader does this). */
int thr_detect_stream(thr_detector *det, const uint8_t *stream, int64_t n_stream_bytes, int64_t first_block,
                      thr_record *out, int64_t *n_blocks);
int thr_detect_stream_device(thr_detector *det, const uint8_t *d_stream, int64_t n_stream_bytes,
                             const int64_t *d_block_idx, int32_t n_blocks, thr_record *d_out);

/* ---- detection, device-resident buffers (async on the handle's stream) ----
 * The same call as thr_detect_batch (thrifty/detect.py:60-78 per block) for callers that keep the samples on the GPU;
 * no counterpart in the reference.  n_blocks <= max_batch.  Pointers are device pointers. */
int thr_detect_batch_device(thr_detector *det, const uint8_t *d_raw, const int64_t *d_block_idx,
                            int32_t n_blocks, thr_record *d_out);
int thr_detect_batch_device_c64(thr_detector *det, const float *d_iq, const int64_t *d_block_idx,
                                int32_t n_blocks, thr_record *d_out);

/* One block with the intermediate arrays of Detector(yield_data=True) (detect.py:75-76):
 * shifted_fft: N complex64 (may be NULL), corr: (N-L+1) complex64 (may be NULL),
 * fft_mag: N float32 |FFT(block)| (may be NULL).  Host pointers; template 0. */
int thr_detect_block_data(thr_detector *det, const uint8_t *raw, const float *iq, int64_t block_idx,
                          thr_record *out, float *shifted_fft, float *corr, float *fft_mag);

/* ---- the stage boundary of the reference's plug-in seam ----
 * The reference's other callers use the two halves of Detector.detect on their own
 * (scripts/chip_rate_search.py:44-55,121-127, thrifty/template_extract.py:36-58):
 *   thr_sync_batch  == thrifty/carrier_sync.py:52-76,82-118 DefaultSynchronizer(thresh, window, block_len, carrier_len)(block)
 *                      -> (shifted_fft, CarrierSyncInfo): carrier decision, Dirichlet fit, mix, FFT#2.  One record per
 *                      block with the carrier fields filled (flags & THR_FLAG_CARRIER_DETECTED; corr fields NaN / -1) and
 *                      the shifted spectrum, N complex64 in natural bin order per block (all zeros where no carrier was
 *                      found: the reference returns None there).  raw (uint8 I/Q) or iq (complex64), the other NULL.
 *   thr_soa_batch   == thrifty/soa_estimator.py:42-92 SoaEstimator(template, thresh, block_len, history_len)(fft)
 *                      -> (detected, CorrDetectionInfo, corr): x conj(FFT(template)), IFFT, windowed peak, noise,
 *                      threshold, Gaussian interpolation.  fft: n_blocks * N complex64 shifted spectra (host); one record
 *                      per block with the corr fields filled (flags & THR_FLAG_CORR_DETECTED, soa = (N-H) block_idx +
 *                      sample + offset) and, if corr != NULL, the correlation c[0 .. N-L] as complex64 per block.
 * Both run the same kernel code as thr_detect_batch, cut at the boundary; one-template detectors. */
int thr_sync_batch(thr_detector *det, const uint8_t *raw, const float *iq, const int64_t *block_idx, int64_t n_blocks,
                   thr_record *out, float *shifted_fft);
int thr_soa_batch(thr_detector *det, const float *fft, const int64_t *block_idx, int64_t n_blocks, thr_record *out,
                  float *corr);

/* ---- several GPUs behind one handle (thrifty/detect.py:217-223: the loop over blocks is the seam) ----
 * The detect path has no cross-block state, so a batch shards into contiguous pieces with no data-path collective: one
 * thr_detector + one host thread per device.  thr_group_detect_batch hands out chunks of the batch from one counter (a GPU
 * behind a slower PCIe path takes fewer), raw streams are cut into one stripe per GPU with an H-sample halo at each stripe
 * start, `.card` text into one stripe per GPU at line boundaries.  Every piece's records land at its blocks' positions in
 * the caller's array, so `out` is in input order exactly as from one GPU (and byte-identical to it).  `devices` are CUDA ordinals; the config's own `device` field is ignored.  Worker threads pin themselves to
 * the CPUs of their GPU's NUMA node when sysfs exposes it; thr_group_host_alloc returns page-locked memory of
 * n_devices * bytes_per_device bytes whose g-th part lives on the g-th GPU's node (part size rounded up to the page size:
 * use thr_group_size() * bytes_per_device only when bytes_per_device is a multiple of 4096). */
typedef struct thr_group thr_group;
int  thr_group_create(const thr_config *cfg, const int32_t *devices, int32_t n_devices, thr_group **out);
void thr_group_destroy(thr_group *grp);
const char *thr_group_last_error(const thr_group *grp);     /* grp == NULL: last failed thr_group_create */
int  thr_group_size(const thr_group *grp);
thr_detector *thr_group_member(thr_group *grp, int32_t i);  /* the i-th device's handle (thr_get_info, ...) */
int  thr_group_numa_node(const thr_group *grp, int32_t i, int32_t *thread_bound);   /* -1: unknown */
int  thr_group_detect_batch(thr_group *grp, const uint8_t *raw, const int64_t *block_idx, int64_t n_blocks,
                            thr_record *out);
int  thr_group_detect_stream(thr_group *grp, const uint8_t *stream, int64_t n_stream_bytes, int64_t first_block,
                             thr_record *out, int64_t *n_blocks);
int  thr_group_detect_card(thr_group *grp, const char *text, size_t len, int32_t final_chunk, int64_t max_blocks,
                           double *timestamps, int64_t *block_idx, thr_record *out, int64_t *n_blocks,
                           int64_t *consumed);
void *thr_group_host_alloc(thr_group *grp, size_t bytes_per_device);
void  thr_group_host_free(thr_group *grp, void *p, size_t bytes_per_device);

/* ---- `identify`: the step after detect (thrifty/identify.py:26-166) on the GPU ----
 * Columns of detections (one value per detection; any order, any number of receivers), host pointers in and out:
 *   thr_identify_classify       identify.py:106-118 classify_transmitters: txid = the last frequency-map range
 *                               (map_rxid, map_txid, [map_start, map_stop] in bins, receiver offset already added) that
 *                               contains carrier_bin + carrier_offset, else -1
 *   thr_identify_bin_histogram  identify.py:39-41: carrier-bin histogram of one receiver (first_bin = its smallest bin);
 *                               counts == NULL only queries first_bin / n_bins.  The caller scans the counts for peaks
 *                               (identify.py:43-61, a sequential pass over ~100 bins) and passes the edges to
 *   thr_identify_digitize       identify.py:98-99: txid = np.digitize(carrier_bin, edges) - 1 for that receiver's rows
 *   thr_identify_duplicates     identify.py:134-164 identify_duplicates: keep[i] = 0 for unidentified detections and for
 *                               the weaker of two detections in adjacent blocks that are neighbours in (rxid, txid, block,
 *                               timestamp) order (cyclic neighbours, as np.roll; `block` is int32 as in toads_array)
 * Errors: negative status, message from thr_identify_last_error(). */
int thr_identify_classify(int32_t device, int64_t n, const int32_t *rxid, const int32_t *carrier_bin,
                          const double *carrier_offset, int32_t n_map, const int32_t *map_rxid, const int32_t *map_txid,
                          const double *map_start, const double *map_stop, int32_t *txid_out);
int thr_identify_bin_histogram(int32_t device, int64_t n, const int32_t *rxid, const int32_t *carrier_bin,
                               int32_t which_rxid, int32_t *first_bin, int32_t *n_bins, uint32_t *counts,
                               int32_t counts_cap);
int thr_identify_digitize(int32_t device, int64_t n, const int32_t *rxid, const int32_t *carrier_bin, int32_t which_rxid,
                          int32_t n_edges, const int64_t *edges, int32_t *txid_inout);
int thr_identify_duplicates(int32_t device, int64_t n, const int32_t *rxid, const int32_t *txid, const int32_t *block,
                            const double *timestamp, const double *energy, uint8_t *keep_out);
const char *thr_identify_last_error(void);

/* ---- .toad text (thrifty/toads_data.py:47-61 DetectionResult.serialize; detect.py:218-219 writes detected blocks only) ----
 * Formats the records with THR_FLAG_CORR_DETECTED among recs[0], recs[stride], ... (n blocks; stride = n_templates) as
 * "{rxid} [{txid} ]{t:.6f} {block} {soa:.8f} {sample} {offset} {energy} {noise} {bin} {offset} {energy} {noise}\n", the
 * free-format fields exactly as Python prints them (shortest round-trip digits, CPython's layout).  txids == NULL writes
 * a .toad; otherwise a .toads line with txids[i] (toads_data.py:57-60).  cap must hold 1024 bytes per detected record;
 * THR_ERR_NOMEM with *used = that size otherwise.  Host-side; several threads for large n. */
int thr_format_toad(const thr_record *recs, const double *timestamps, int64_t n, int64_t stride, int32_t rxid,
                    const int32_t *txids, char *buf, size_t cap, size_t *used);

/* ---- stream / timing plumbing (no counterpart in the reference) ---- */
int thr_set_stream(thr_detector *det, void *cuda_stream);   /* NULL -> handle's own stream */
int thr_synchronize(thr_detector *det);
int thr_timer_start(thr_detector *det);                     /* CUDA event on the handle's stream */
int thr_timer_stop(thr_detector *det, float *elapsed_ms);   /* records, synchronises, returns ms  */

/* ---- memory helpers (so callers need no CUDA binding of their own; no counterpart in the reference) ---- */
void *thr_host_alloc(size_t bytes);                         /* pinned host memory */
void  thr_host_free(void *p);
void *thr_device_alloc(int device, size_t bytes);
void  thr_device_free(int device, void *p);
int   thr_memcpy_h2d(int device, void *dst, const void *src, size_t bytes);
int   thr_memcpy_d2h(int device, void *dst, const void *src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* THRIFTY_B200_H */
