// fastdet_ref_driver.cpp -- TEST INFRASTRUCTURE (oracle).  Drives the reference's OWN native sources
// (compiled in place from /root/reference by oracle/Makefile; nothing is copied into this repo):
//
//   fastcard/fastcard.c reader.c raw_reader.c card_reader.c lib/base64.c rawconv.c cardet.c fft.c
//   fastdet/corr_detector.cpp fastcard_wrappers.cpp
//
// through the same calls as the reference's main loop (fastdet/fastdet.cpp:113-186): a
// CarrierDetector over a .card or raw file, a CorrDetector on every carrier-positive block, and the
// SoA assembly of fastdet.cpp:184-186.  What is NOT the reference here: the FFTW3f and VOLK entry
// points (oracle/shim/, a double-precision radix-2 FFT and VOLK's generic scalar loops) and the
// librtlsdr reader (stubbed; hardware capture is out of scope).  Output: one record per block read,
// carrier-less blocks included, so that the CUDA `fastdet` mode can be compared block by block.
//
// Only tests/, oracle/make_golden_fastdet.py and bench.py's CPU legs may load the resulting library.
#include <cstdio>
#include <cstring>
#include <exception>
#include <vector>

#include "fastcard_wrappers.h"     // /root/reference/fastdet
#include "corr_detector.h"         // /root/reference/fastdet
#include <fastcard/rtlsdr_reader.h>

extern "C" {

// librtlsdr-backed reader: not built (needs hardware + librtlsdr); fastcard.c references the symbols.
reader_t *rtlsdr_reader_new(reader_settings_t, rtlsdr_settings_t *) { return NULL; }
void rtlsdr_reader_print_histogram(reader_t *, FILE *) {}

struct ref_fastdet_record {
    int64_t block_idx;
    int64_t ts_sec, ts_usec;
    double soa;                 // fastdet.cpp:184-186
    double corr_offset;         // CorrDetection::peak_offset
    double carrier_offset;      // CorrDetection::carrier_offset (parabolic, reporting only)
    float carrier_max, carrier_noise, carrier_threshold, fft_sum;   // cardet_detection_t (powers)
    float corr_peak_power, corr_noise_power, corr_threshold;        // CorrDetection (powers)
    int32_t carrier_detected, carrier_argmax;
    int32_t corr_detected, corr_peak_idx;
    int32_t pad;
};

// Returns the number of blocks read (records written, capped at max_out), or <0 on error.
int ref_fastdet_run(const char *input_path, int input_card, int block_len, int history_len,
                    float thresh_const, float thresh_snr, int win_min, int win_max,
                    const float *tpl, int tpl_len, float corr_thresh_const, float corr_thresh_snr,
                    ref_fastdet_record *out, int max_out, char *errbuf, int errbuf_len) {
    try {
        fargs_t args;
        std::memset(&args, 0, sizeof args);
        args.block_len = block_len;
        args.history_len = history_len;
        args.threshold_const = thresh_const;
        args.threshold_snr = thresh_snr;
        args.carrier_freq_min = win_min;
        args.carrier_freq_max = win_max;
        args.skip = 0;
        args.input_file = input_path;
        args.wisdom_file = NULL;
        args.input_card = input_card != 0;
        args.silent = true;

        CarrierDetector carrier_det(&args);
        std::vector<float> template_samples(tpl, tpl + tpl_len);
        CorrDetector corr_detect(template_samples, block_len, history_len, corr_thresh_const, corr_thresh_snr);
        carrier_det.start();
        int n = 0;
        while (carrier_det.process_next()) {                       // fastdet.cpp:163
            const fastcard_data_t &carrier = carrier_det.data();
            if (n < max_out) {
                ref_fastdet_record &r = out[n];
                std::memset(&r, 0, sizeof r);
                r.block_idx = carrier.block->index;
                r.ts_sec = carrier.block->timestamp.tv_sec;
                r.ts_usec = carrier.block->timestamp.tv_usec;
                r.carrier_detected = carrier.detected ? 1 : 0;
                if (carrier.detected) {                            // fastdet.cpp:175-186
                    const CorrDetection corr = corr_detect.detect(carrier);
                    r.carrier_argmax = carrier.detection.argmax;
                    r.carrier_max = carrier.detection.max;
                    r.carrier_noise = carrier.detection.noise;
                    r.carrier_threshold = carrier.detection.threshold;
                    r.fft_sum = carrier.detection.fft_sum;
                    r.corr_detected = corr.detected ? 1 : 0;
                    r.corr_peak_idx = corr.peak_idx;
                    r.corr_offset = corr.peak_offset;
                    r.corr_peak_power = corr.peak_power;
                    r.corr_noise_power = corr.noise_power;
                    r.corr_threshold = corr.threshold;
                    r.carrier_offset = corr.carrier_offset;
                    r.soa = ((double)(args.block_len - args.history_len) * (double)carrier.block->index +
                             corr.peak_idx) + corr.peak_offset;
                }
            }
            ++n;
        }
        return n;
    } catch (std::exception &e) {
        if (errbuf && errbuf_len > 0) std::snprintf(errbuf, errbuf_len, "%s", e.what());
        return -1;
    }
}

// rawconv LUT of the reference (fastcard/rawconv.c:5-28) for the byte pair (i, q)
void ref_rawconv(const uint8_t *raw, int n_samples, float *out_iq) {
    static rawconv_t *lut = NULL;
    if (!lut) {
        lut = (rawconv_t *)std::malloc(sizeof(rawconv_t));
        rawconv_init(lut);
    }
    rawconv_to_complex(lut, (fcomplex *)out_iq, (uint16_t *)raw, n_samples);
}

}  // extern "C"
