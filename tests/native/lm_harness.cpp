// lm_harness.cpp -- TEST INFRASTRUCTURE: compiles thrifty_b200/csrc/dirichlet_lm.cuh for the host so that the
// device's Levenberg-Marquardt code can be compared with scipy.optimize.leastsq on the CPU (tests/test_lm_fit.py).
// Not part of the product: nothing under thrifty_b200/ loads this.
#include "../../thrifty_b200/csrc/dirichlet_lm.cuh"

extern "C" int lm_fit_host(const double *y, double W, double N, double *amplitude, double *offset, int *nfev) {
    const double piW = 3.141592653589793 * W;
    auto weights = [&](double da, double db, double *ga, double *gb) {
        for (int i = 0; i < thr::lm::M; ++i) {
            ga[i] = thr::lm::weight((double)(i - 3) - da, piW, N, W);
            gb[i] = thr::lm::weight((double)(i - 3) - db, piW, N, W);
        }
    };
    thr::lm::Rows rows;
    for (int i = 0; i < thr::lm::M; ++i) rows.y[i] = y[i];
    const thr::lm::Result r = thr::lm::fit(thr::lm::SerialExec(), weights, rows, y[3], 0.0);
    *amplitude = r.amplitude;
    *offset = r.offset;
    *nfev = r.nfev;
    return r.info;
}

// the short cut (quick_fit): returns 1 if its result may stand in for lmdif's
extern "C" int lm_quick_host(const double *y, double W, double N, double *amplitude, double *offset, double *slack) {
    const double piW = 3.141592653589793 * W;
    auto derivs_d = [&](double d, double *D, double *Dz) {
        for (int i = 0; i < thr::lm::M; ++i) thr::lm::kernel_deriv<double>((double)(i - 3) - d, piW, N, W, D[i], Dz[i]);
    };
    auto derivs_f = [&](float d, float *D, float *Dz) {
        for (int i = 0; i < thr::lm::M; ++i)
            thr::lm::kernel_deriv<float>((float)(i - 3) - d, (float)piW, (float)N, (float)W, D[i], Dz[i]);
    };
    thr::lm::Rows rows;
    for (int i = 0; i < thr::lm::M; ++i) rows.y[i] = y[i];
    const thr::lm::Quick q = thr::lm::quick_fit(thr::lm::SerialExec(), derivs_f, derivs_d, rows, N / W);
    *amplitude = q.amplitude;
    *offset = q.offset;
    *slack = q.slack;
    return q.ok ? 1 : 0;
}
