// dirichlet_lm.cuh -- sub-bin carrier offset: Levenberg-Marquardt fit of a Dirichlet kernel, float64.
//
// The reference obtains the offset with scipy.optimize.curve_fit(_fit_model, xdata, ydata, p0) on the
// 7 spectrum magnitudes around the carrier peak (thrifty/carrier_sync.py:150-196, model :121-132).
// curve_fit (method 'lm') is MINPACK's lmdif: forward-difference Jacobian (step sqrt(eps_machine) |x|),
// QR with column pivoting, Levenberg-Marquardt parameter by More's lmpar, trust region `delta`, column
// scaling diag = max column norm seen so far (mode 1), factor = 100, ftol = xtol = 1.49012e-8, gtol = 0,
// maxfev = 200 (n + 1).  When the Dirichlet main lobe is much wider than the 7 fitted bins the cost surface
// is nearly flat along the offset and WHERE the iteration stops decides the 4th digit of the answer, so
// the device fit follows the same algorithm, in float64, with the same stopping tests; it then stops at the
// same iterate as the reference.  This file restates the published MINPACK algorithm (lmdif / fdjac2 /
// qrfac / lmpar / qrsolv, Argonne National Laboratory, 1980) for n = 2 parameters and m = 7 points.
//
// The code is plain scalar C++ (host + device) so that tests/native/lm_harness.cpp can run the very same
// source on the CPU against scipy (tests/test_lm_fit.py: identical iterates, identical nfev); only the evaluation
// of the model weights is supplied by the caller (on the device: one lane per point, see detect_kernel.cuh).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define THR_HD __host__ __device__ __forceinline__
#else
#define THR_HD inline
#endif
// Loops around divisions / square roots / the control flow of the iteration stay rolled on the device: a float64
// division expands to ~25 instructions, and the code runs on one warp whose instruction footprint should stay small.
// Loops over the two parameters are unrolled and the column permutation of the pivoted QR is a single flag, so that
// every small array is indexed statically and stays out of local memory.
#if defined(__CUDA_ARCH__)
#define THR_ROLLED _Pragma("unroll 1")
#define THR_UNROLLED _Pragma("unroll")
#else
#define THR_ROLLED
#define THR_UNROLLED
#endif

namespace thr {
namespace lm {

constexpr int M = 7;                       // points (bins k-3 .. k+3)
constexpr int NP = 2;                      // parameters (amplitude, offset)
constexpr double EPSMCH = 2.220446049250313e-16;
constexpr double DWARF = 2.2250738585072014e-308;
constexpr double TOL = 1.49012e-8;         // scipy.optimize.leastsq default ftol = xtol
constexpr double FACTOR = 100.0;
constexpr int MAXFEV = 200 * (NP + 1);

// Weight of the reference's model at z = x - offset: |sin(pi W z / N) / sin(pi z / N) / W|, evaluated in the
// reference's operation order (carrier_sync.py:129-131, :180-183); 0/0 -> 1.  A residual is amplitude * weight - y.
THR_HD double weight(double z, double piW, double N, double W) {
    const double s2 = sin((3.141592653589793 * z) / N);
    double w = sin((piW * z) / N) / s2 / W;
    if (w != w) w = 1.0;
    return fabs(w);
}

THR_HD double enorm2(double a, double b) { return sqrt(a * a + b * b); }

// ipvt of MINPACK for n = 2 is either (0, 1) or (1, 0): `swp` says which; pv(v, j, swp) = v[ipvt[j]]
THR_HD double pv(const double (&v)[NP], int j, bool swp) { return ((j != 0) != swp) ? v[1] : v[0]; }
THR_HD void pset(double (&v)[NP], int j, bool swp, double val) {
    if ((j != 0) != swp) v[1] = val; else v[0] = val;
}

// qrsolv for n = 2: r = upper triangle R (r[i][j], i <= j) of the pivoted QR, overwritten below the diagonal
// with the transposed strict upper triangle of S; returns x (solution of the damped system) and sdiag.
THR_HD void qrsolv2(double (&r)[NP][NP], bool swp, const double (&diag)[NP], const double (&qtb)[NP],
                    double (&x)[NP], double (&sdiag)[NP]) {
    double wa[NP];
    THR_UNROLLED
    for (int j = 0; j < NP; ++j) {
        THR_UNROLLED
        for (int i = j; i < NP; ++i) r[i][j] = r[j][i];
        x[j] = r[j][j];
        wa[j] = qtb[j];
    }
    THR_UNROLLED
    for (int j = 0; j < NP; ++j) {
        const double dl = pv(diag, j, swp);
        if (dl != 0.0) {
            THR_UNROLLED
            for (int k = j; k < NP; ++k) sdiag[k] = 0.0;
            sdiag[j] = dl;
            double qtbpj = 0.0;
            THR_UNROLLED
            for (int k = j; k < NP; ++k) {
                if (sdiag[k] == 0.0) continue;
                double cs, sn;
                if (fabs(r[k][k]) < fabs(sdiag[k])) {
                    const double cotan = r[k][k] / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    cs = sn * cotan;
                } else {
                    const double tn = sdiag[k] / r[k][k];
                    cs = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    sn = cs * tn;
                }
                r[k][k] = cs * r[k][k] + sn * sdiag[k];
                const double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                THR_UNROLLED
                for (int i = k + 1; i < NP; ++i) {
                    const double t2 = cs * r[i][k] + sn * sdiag[i];
                    sdiag[i] = -sn * r[i][k] + cs * sdiag[i];
                    r[i][k] = t2;
                }
            }
        }
        sdiag[j] = r[j][j];
        r[j][j] = x[j];
    }
    // triangular solve; a zero on the diagonal of S truncates the system (least-squares solution)
    const bool s0 = sdiag[0] == 0.0, s1 = sdiag[1] == 0.0;
    if (s0) {
        wa[0] = 0.0;
        wa[1] = 0.0;
    } else if (s1) {
        wa[1] = 0.0;
        wa[0] = wa[0] / sdiag[0];
    } else {
        wa[1] = wa[1] / sdiag[1];
        wa[0] = (wa[0] - r[1][0] * wa[1]) / sdiag[0];
    }
    THR_UNROLLED
    for (int j = 0; j < NP; ++j) pset(x, j, swp, wa[j]);
}

// lmpar for n = 2: Levenberg-Marquardt parameter `par` such that ||diag x|| is within 10 % of delta.
THR_HD void lmpar2(double (&r)[NP][NP], bool swp, const double (&diag)[NP], const double (&qtb)[NP], double delta,
                   double &par, double (&x)[NP], double (&sdiag)[NP]) {
    double wa1[NP], wa2[NP];
    // Gauss-Newton direction (rank-deficient R: least-squares solution)
    const bool z0 = r[0][0] == 0.0, z1 = r[1][1] == 0.0;
    const bool full_rank = !z0 && !z1;
    if (z0) {
        wa1[0] = 0.0;
        wa1[1] = 0.0;
    } else if (z1) {
        wa1[1] = 0.0;
        wa1[0] = qtb[0] / r[0][0];
    } else {
        wa1[1] = qtb[1] / r[1][1];
        wa1[0] = (qtb[0] - r[0][1] * wa1[1]) / r[0][0];
    }
    THR_UNROLLED
    for (int j = 0; j < NP; ++j) pset(x, j, swp, wa1[j]);
    int iter = 0;
    THR_UNROLLED
    for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = enorm2(wa2[0], wa2[1]);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {
        par = 0.0;
        return;
    }
    double parl = 0.0;
    if (full_rank) {
        THR_UNROLLED
        for (int j = 0; j < NP; ++j) wa1[j] = pv(diag, j, swp) * (pv(wa2, j, swp) / dxnorm);
        wa1[0] = wa1[0] / r[0][0];
        wa1[1] = (wa1[1] - r[0][1] * wa1[0]) / r[1][1];
        const double temp = enorm2(wa1[0], wa1[1]);
        parl = ((fp / delta) / temp) / temp;
    }
    wa1[0] = (r[0][0] * qtb[0]) / pv(diag, 0, swp);
    wa1[1] = (r[0][1] * qtb[0] + r[1][1] * qtb[1]) / pv(diag, 1, swp);
    const double gnorm = enorm2(wa1[0], wa1[1]);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = DWARF / fmin(delta, 0.1);
    par = fmax(par, parl);
    par = fmin(par, paru);
    if (par == 0.0) par = gnorm / dxnorm;
    THR_ROLLED
    for (;;) {
        ++iter;
        if (par == 0.0) par = fmax(DWARF, 0.001 * paru);
        double temp = sqrt(par);
        THR_UNROLLED
        for (int j = 0; j < NP; ++j) wa1[j] = temp * diag[j];
        qrsolv2(r, swp, wa1, qtb, x, sdiag);
        THR_UNROLLED
        for (int j = 0; j < NP; ++j) wa2[j] = diag[j] * x[j];
        dxnorm = enorm2(wa2[0], wa2[1]);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        THR_UNROLLED
        for (int j = 0; j < NP; ++j) wa1[j] = pv(diag, j, swp) * (pv(wa2, j, swp) / dxnorm);
        wa1[0] = wa1[0] / sdiag[0];
        wa1[1] = (wa1[1] - r[1][0] * wa1[0]) / sdiag[1];
        temp = enorm2(wa1[0], wa1[1]);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, par);
        if (fp < 0.0) paru = fmin(paru, par);
        par = fmax(parl, par + parc);
    }
}

struct Result {
    double amplitude, offset;
    int nfev, info;
};

// Row data of the 7-point problem.  On the host one thread owns all of it; on the device it lives in shared memory and
// the lanes of one warp split the element-wise work (lane r writes row r, then the warp synchronises) while every lane
// carries the scalar state of the iteration redundantly and forms the sums over rows itself, in row order, from the
// shared rows.  Slot 7 of each array is a dummy (the lane next to the last row evaluates a duplicate of row 6).
struct Rows {
    double y[8];                 // the magnitudes being fitted
    double fvec[8], wa4[8];      // residuals at x / at the trial point (also scratch for Q^T fvec)
    double a[NP][8];             // Jacobian columns, then Householder vectors
    double gx[8], gh[8];         // model weights at x and at x + h e_offset
    double gt[8], gth[8];        // the same at the trial point
};

// Execution policy of the host build: one thread, all rows.
struct SerialExec {
    template <class F>
    THR_HD void each(int from, F &&f) const {
        for (int r = from; r < M; ++r) f(r);
    }
    THR_HD void sync() const {}
};

// lmdif for m = 7, n = 2.  weights(da, db, ga, gb) must fill rows 0..6 of ga / gb with the model weights at the offsets
// da / db and synchronise.  The amplitude enters the model linearly, so the forward-difference column of the amplitude
// re-uses the weights at the current offset, and the offset column of the NEXT iterate (offset + h) is evaluated
// together with the trial point it belongs to: one round of sines per trial step, with exactly the values lmdif /
// fdjac2 would compute.  `ex.each(from, f)` runs f(r) for the rows r >= from (device: one lane per row),
// `ex.sync()` orders the rows' writes before everybody's reads (device: __syncwarp()).
template <class Exec, class Weights>
THR_HD Result fit(Exec &&ex, Weights &&weights, Rows &w, double a0, double d0) {
    double x[NP] = {a0, d0};
    double diag[NP], qtf[NP], wa1[NP], wa2[NP], wa3[NP], rdiag[NP], acnorm[NP];
    bool swp = false;                        // column permutation of the QR: (0, 1) or (1, 0)
    Result res;
    res.info = 0;
    int nfev = 1;
    const double eps = sqrt(EPSMCH);
    auto step_of = [&](double v) {
        const double h = eps * fabs(v);
        return h == 0.0 ? eps : h;
    };
    weights(x[1], x[1] + step_of(x[1]), w.gx, w.gh);
    ex.each(0, [&](int r) { w.fvec[r] = x[0] * w.gx[r] - w.y[r]; });
    ex.sync();
    double fnorm = 0.0;
    for (int i = 0; i < M; ++i) fnorm += w.fvec[i] * w.fvec[i];
    fnorm = sqrt(fnorm);
    double par = 0.0, delta = 0.0, xnorm = 0.0;
    int iter = 1;
    THR_ROLLED
    for (;;) {
        // fdjac2: forward differences (x[j] + h is formed first, then (f(x + h e_j) - f(x)) / h)
        {
            const double h0 = step_of(x[0]), ah = x[0] + h0;
            const double h1 = step_of(x[1]);
            ex.each(0, [&](int r) {
                w.a[0][r] = ((ah * w.gx[r] - w.y[r]) - w.fvec[r]) / h0;
                w.a[1][r] = ((x[0] * w.gh[r] - w.y[r]) - w.fvec[r]) / h1;
            });
            ex.sync();
        }
        nfev += NP;
        // qrfac with column pivoting
        THR_UNROLLED
        for (int j = 0; j < NP; ++j) {
            double s2 = 0.0;
            for (int i = 0; i < M; ++i) s2 += w.a[j][i] * w.a[j][i];
            acnorm[j] = sqrt(s2);
            rdiag[j] = acnorm[j];
            wa3[j] = rdiag[j];
        }
        swp = rdiag[1] > rdiag[0];
        if (swp) {                           // bring the column of largest norm into the pivot position
            ex.sync();
            ex.each(0, [&](int r) {
                const double t = w.a[0][r];
                w.a[0][r] = w.a[1][r];
                w.a[1][r] = t;
            });
            ex.sync();
            rdiag[1] = rdiag[0];
            wa3[1] = wa3[0];
        }
        THR_UNROLLED
        for (int j = 0; j < NP; ++j) {
            double ajnorm = 0.0;
            for (int i = j; i < M; ++i) ajnorm += w.a[j][i] * w.a[j][i];
            ajnorm = sqrt(ajnorm);
            if (ajnorm != 0.0) {
                if (w.a[j][j] < 0.0) ajnorm = -ajnorm;
                ex.sync();
                ex.each(j, [&](int r) {
                    double v = w.a[j][r] / ajnorm;
                    if (r == j) v += 1.0;
                    w.a[j][r] = v;
                });
                ex.sync();
                THR_UNROLLED
                for (int k = j + 1; k < NP; ++k) {
                    double sum = 0.0;
                    for (int i = j; i < M; ++i) sum += w.a[j][i] * w.a[k][i];
                    const double temp = sum / w.a[j][j];
                    ex.sync();
                    ex.each(j, [&](int r) { w.a[k][r] -= temp * w.a[j][r]; });
                    ex.sync();
                    if (rdiag[k] != 0.0) {
                        const double t = w.a[k][j] / rdiag[k];
                        rdiag[k] *= sqrt(fmax(0.0, 1.0 - t * t));
                        const double q = rdiag[k] / wa3[k];
                        if (0.05 * (q * q) <= EPSMCH) {
                            double s2 = 0.0;
                            for (int i = j + 1; i < M; ++i) s2 += w.a[k][i] * w.a[k][i];
                            rdiag[k] = sqrt(s2);
                            wa3[k] = rdiag[k];
                        }
                    }
                }
            }
            rdiag[j] = -ajnorm;
        }
        if (iter == 1) {
            THR_UNROLLED
            for (int j = 0; j < NP; ++j) {
                diag[j] = acnorm[j];
                if (acnorm[j] == 0.0) diag[j] = 1.0;
            }
            xnorm = enorm2(diag[0] * x[0], diag[1] * x[1]);
            delta = FACTOR * xnorm;
            if (delta == 0.0) delta = FACTOR;
        }
        // qtf = first n components of Q^T fvec; R = (rdiag on the diagonal, a[1][0] above it)
        ex.each(0, [&](int r) { w.wa4[r] = w.fvec[r]; });
        ex.sync();
        THR_UNROLLED
        for (int j = 0; j < NP; ++j) {
            if (w.a[j][j] != 0.0) {
                double sum = 0.0;
                for (int i = j; i < M; ++i) sum += w.a[j][i] * w.wa4[i];
                const double temp = -sum / w.a[j][j];
                ex.sync();
                ex.each(j, [&](int r) { w.wa4[r] += w.a[j][r] * temp; });
                ex.sync();
            }
            qtf[j] = w.wa4[j];
        }
        double r[NP][NP];                    // r[i][j], i <= j: upper triangle of R
        r[0][0] = rdiag[0];
        r[0][1] = w.a[1][0];
        r[1][1] = rdiag[1];
        r[1][0] = 0.0;
        // norm of the scaled gradient (gtol = 0: only the epsmch test further down can fire)
        double gnorm = 0.0;
        if (fnorm != 0.0) {
            THR_UNROLLED
            for (int j = 0; j < NP; ++j) {
                const double acl = pv(acnorm, j, swp);
                if (acl != 0.0) {
                    double sum = 0.0;
                    THR_UNROLLED
                    for (int i = 0; i <= j; ++i) sum += r[i][j] * (qtf[i] / fnorm);
                    gnorm = fmax(gnorm, fabs(sum / acl));
                }
            }
        }
        if (gnorm <= 0.0) {                  // gtol = 0
            res.info = 4;
            break;
        }
        THR_UNROLLED
        for (int j = 0; j < NP; ++j) diag[j] = fmax(diag[j], acnorm[j]);
        // inner loop: trial steps until one is accepted
        double ratio;
        THR_ROLLED
        do {
            double sdiag[NP];
            lmpar2(r, swp, diag, qtf, delta, par, wa1, sdiag);
            THR_UNROLLED
            for (int j = 0; j < NP; ++j) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            const double pnorm = enorm2(wa3[0], wa3[1]);
            if (iter == 1) delta = fmin(delta, pnorm);
            ex.sync();
            weights(wa2[1], wa2[1] + step_of(wa2[1]), w.gt, w.gth);
            ++nfev;
            ex.each(0, [&](int rr) { w.wa4[rr] = wa2[0] * w.gt[rr] - w.y[rr]; });
            ex.sync();
            double fnorm1 = 0.0;
            for (int i = 0; i < M; ++i) fnorm1 += w.wa4[i] * w.wa4[i];
            fnorm1 = sqrt(fnorm1);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) {
                const double q = fnorm1 / fnorm;
                actred = 1.0 - q * q;
            }
            // predicted reduction and directional derivative: R (P^T p)
            double w3[NP] = {0.0, 0.0};
            THR_UNROLLED
            for (int j = 0; j < NP; ++j) {
                const double temp = pv(wa1, j, swp);
                THR_UNROLLED
                for (int i = 0; i <= j; ++i) w3[i] += r[i][j] * temp;
            }
            const double temp1 = enorm2(w3[0], w3[1]) / fnorm;
            const double temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp = 0.5;
                if (actred < 0.0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par = par / temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par = 0.5 * par;
            }
            if (ratio >= 1e-4) {
                x[0] = wa2[0];
                x[1] = wa2[1];
                ex.sync();
                ex.each(0, [&](int rr) {
                    w.fvec[rr] = w.wa4[rr];
                    w.gx[rr] = w.gt[rr];
                    w.gh[rr] = w.gth[rr];
                });
                ex.sync();
                xnorm = enorm2(diag[0] * x[0], diag[1] * x[1]);
                fnorm = fnorm1;
                ++iter;
            }
            if (fabs(actred) <= TOL && prered <= TOL && 0.5 * ratio <= 1.0) res.info = 1;
            if (delta <= TOL * xnorm) res.info = 2;
            if (fabs(actred) <= TOL && prered <= TOL && 0.5 * ratio <= 1.0 && res.info == 2) res.info = 3;
            if (res.info != 0) break;
            if (nfev >= MAXFEV) res.info = 5;
            if (fabs(actred) <= EPSMCH && prered <= EPSMCH && 0.5 * ratio <= 1.0) res.info = 6;
            if (delta <= EPSMCH * xnorm) res.info = 7;
            if (gnorm <= EPSMCH) res.info = 8;
            if (res.info != 0) break;
        } while (ratio < 1e-4);
        if (res.info != 0) break;
    }
    res.amplitude = x[0];
    res.offset = x[1];
    res.nfev = nfev;
    return res;
}

// ---------------------------------------------------------------------------------------------------------------
// Short cut.  lmdif's answer is the iterate at which its stopping tests fire, and those fire within
//     slack = sqrt(2 ftol C cov_dd)      (C = residual sum of squares, cov_dd = [(J^T J)^-1]_dd at the minimum)
// of the least-squares minimum along the offset: beyond that distance one more step still lowers the cost by more
// than ftol C.  (Measured on 2 100 fits, well- and ill-conditioned: |offset_lmdif - offset_min| <= 0.18 slack.)
// Where the slack is far below the parity bar of 1e-4 bins the minimum itself is therefore the answer, and a plain
// Gauss-Newton iteration on the normal equations finds it in a fraction of the instructions.  quick_fit() does that and
// reports whether its result may stand in for lmdif's: converged, every step lowered the cost (no damping: nothing
// path dependent happened), |offset| < 1 and slack < QUICK_SLACK.  Otherwise the caller runs fit() -- after restoring
// rows.y, which the short cut's workspace overlays.
constexpr double QUICK_SLACK = 3e-5;
constexpr int QUICK_MAXIT = 10;
constexpr double QUICK_STEP = 1e-4;       // double phase: a step below this ends the iteration (what is left after a step
                                           // is <= 0.1 of it here; the float phase usually leaves one such step of ~1e-5)
constexpr double START_NULL = 1e-3;        // |D| below this at an off-peak point of the starting guess: on a null
constexpr double QUICK_MAX_LOBE = 24.0;    // short cut only for N / W <= 24 bins (beyond, the 7 points sit on the flat top
                                           // of the main lobe and lmdif can stop on xtol far outside the ftol slack)
// D(z) = sin(pi W z / N) / (W sin(pi z / N)) and dD/dz at z = x - offset (carrier_sync.py:121-132); z == 0: (1, 0)
template <class T>
THR_HD void kernel_deriv(T z, T piW, T N, T W, T &D, T &Dz) {
    const T pi = (T)3.141592653589793;
    const T t1 = (piW * z) / N, t2 = (pi * z) / N;
    T s1, c1, s2, c2;
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        sincosf(t1, &s1, &c1);
        sincosf(t2, &s2, &c2);
    } else {
        sincos(t1, &s1, &c1);       // one range reduction per angle
        sincos(t2, &s2, &c2);
    }
#else
    s1 = (T)sin((double)t1), c1 = (T)cos((double)t1), s2 = (T)sin((double)t2), c2 = (T)cos((double)t2);
#endif
    D = s1 / s2 / W;
    Dz = ((piW / N) * c1 * s2 - (pi / N) * s1 * c2) / (s2 * s2) / W;
    if (D != D) {
        D = (T)1;
        Dz = (T)0;
    }
}

// Gauss-Newton on the normal equations of the 7-point problem, in float or double.  Row workspace (shared memory on the
// device: the double-precision lm::Rows is large enough for either instantiation and is reused).
template <class T>
struct GnRows {
    T y[8], D[8], Dz[8], g[8], r[8], jd[8], neg[8];
};
static_assert(sizeof(GnRows<double>) <= sizeof(Rows), "the Gauss-Newton rows live in the lmdif workspace");

template <class T>
struct GnResult {
    T amplitude, offset;
    double slack;           // sqrt(2 ftol C cov_dd) at the last evaluated iterate
    unsigned pattern;       // bit r: D < 0 at point r (sign pattern the iteration settled in)
    bool ok;                // converged, every step lowered the cost, the sign pattern never changed after `settle`
};

// derivs(d, D, Dz) must fill rows 0..6 with the kernel and its z-derivative at the offset d, and synchronise.
// The model |D| has a kink wherever D changes sign (the nulls of the kernel, at z = m N / W), and with a point next to a
// null there is a local minimum on either side of it -- which one an iteration ends in depends on its path.  lmdif starts
// at (y[3], 0) and its first step is the Gauss-Newton step; so the sign pattern of D over the 7 points is recorded after
// the first step from that start (settle = 1) and must not change afterwards -- neither in this run nor in a later one
// that continues from its result (settle = 0, pattern_in = the recorded pattern).  Otherwise the result is not vouched for.
template <class T, class Exec, class Derivs>
THR_HD GnResult<T> gn_fit(Exec &&ex, Derivs &&derivs, GnRows<T> &w, T a0, T d0, T step_tol, int settle, unsigned pattern_in,
                          int max_iter) {
    T A = a0, d = d0;
    GnResult<T> q;
    q.ok = false;
    q.slack = 1.0;
    q.pattern = pattern_in;
    T cost_prev = (T)3e38;
    bool converged = false;
    T saa = 0, sdd = 0, sad = 0, cost = 0;
    THR_ROLLED
    for (int it = 0; it < max_iter; ++it) {
        derivs(d, w.D, w.Dz);
        ex.each(0, [&](int r) {
            const T D = w.D[r];
            const T g = D < (T)0 ? -D : D, gd = D < (T)0 ? w.Dz[r] : -w.Dz[r];   // d|D|/d offset = -sign(D) dD/dz
            w.g[r] = g;
            w.r[r] = A * g - w.y[r];                            // residual
            w.jd[r] = A * gd;                                   // d residual / d offset (d residual / d A = g)
            w.neg[r] = D < (T)0 ? (T)1 : (T)0;
        });
        ex.sync();
        T sar = 0, sdr = 0;
        saa = sdd = sad = cost = 0;
        unsigned pattern = 0;
        for (int i = 0; i < M; ++i) {
            const T g = w.g[i], jd = w.jd[i], r = w.r[i];
            saa += g * g;
            sad += g * jd;
            sdd += jd * jd;
            sar += g * r;
            sdr += jd * r;
            cost += r * r;
            pattern |= (w.neg[i] != (T)0 ? 1u : 0u) << i;
        }
        ex.sync();
        if (it == 0 && settle == 1) {
            // lmdif starts exactly here too, but differentiates one-sidedly; with a point ON a null of the kernel at the
            // start (carrier as long as the block, or half / a third of it) the analytic slope of |D| there is a coin
            // toss and the first steps part ways
            bool on_null = false;
            for (int i = 0; i < M; ++i) on_null = on_null || (i != 3 && w.g[i] < (T)START_NULL);
            if (on_null) break;
        }
        if (it == settle) q.pattern = pattern;
        // (a cost that only moves in its last bits -- the iteration has arrived, in this precision -- is not an increase)
        const T slop = sizeof(T) == 4 ? (T)1.00002 : (T)1.00000000001;
        if (!(cost <= cost_prev * slop) || (it > settle - (settle == 0) && pattern != q.pattern)) break;   // leave it to lmdif
        cost_prev = cost;
        const T det = saa * sdd - sad * sad;
        if (!(det > (T)0)) break;
        const T dA = -(sdd * sar - sad * sdr) / det, dd = -(saa * sdr - sad * sar) / det;
        A += dA;
        d += dd;
        const T adA = dA < (T)0 ? -dA : dA, add = dd < (T)0 ? -dd : dd, aA = A < (T)0 ? -A : A;
        if (it >= settle && add < step_tol && adA <= step_tol * aA) {     // what is left after this step is a small
            converged = true;                                             // fraction of it
            break;
        }
    }
    q.amplitude = A;
    q.offset = d;
    if (converged) {
        const double det = (double)saa * (double)sdd - (double)sad * (double)sad;
        q.slack = sqrt(2.0 * TOL * (double)cost * ((double)saa / det));
        q.ok = true;
    }
    return q;
}

struct Quick {
    double amplitude, offset, slack;
    bool ok;
};

// The short cut in two phases: float Gauss-Newton from lmdif's starting point down to the float noise floor (cheap: the
// FP32 pipe, ~6-digit sines), then the same iteration in double from there -- one or two steps -- to the minimum itself.
// derivs_f / derivs_d: the kernel evaluations in float / double (see gn_fit).
constexpr float QUICK_STEP_F32 = 3e-5f;
template <class Exec, class DerivsF, class DerivsD>
THR_HD Quick quick_fit(Exec &&ex, DerivsF &&derivs_f, DerivsD &&derivs_d, Rows &rows, double lobe /* = N / W */) {
    Quick q;
    q.ok = false;
    q.slack = 1.0;
    q.amplitude = rows.y[3];
    q.offset = 0.0;
    if (!(lobe <= QUICK_MAX_LOBE)) return q;
    const double y0 = rows.y[0], y1 = rows.y[1], y2 = rows.y[2], y3 = rows.y[3], y4 = rows.y[4], y5 = rows.y[5], y6 = rows.y[6];
    ex.sync();
    GnRows<float> &wf = *reinterpret_cast<GnRows<float> *>(&rows);
    ex.each(0, [&](int r) {
        const double yr = r == 0 ? y0 : r == 1 ? y1 : r == 2 ? y2 : r == 3 ? y3 : r == 4 ? y4 : r == 5 ? y5 : y6;
        wf.y[r] = (float)yr;
    });
    ex.sync();
    const GnResult<float> p1 = gn_fit<float>(ex, derivs_f, wf, (float)y3, 0.0f, QUICK_STEP_F32, 1, 0u, QUICK_MAXIT);
    ex.sync();
    GnRows<double> &wd = *reinterpret_cast<GnRows<double> *>(&rows);
    ex.each(0, [&](int r) {
        wd.y[r] = r == 0 ? y0 : r == 1 ? y1 : r == 2 ? y2 : r == 3 ? y3 : r == 4 ? y4 : r == 5 ? y5 : y6;
    });
    ex.sync();
    if (p1.ok) {
        const GnResult<double> p2 = gn_fit<double>(ex, derivs_d, wd, (double)p1.amplitude, (double)p1.offset, QUICK_STEP, 0,
                                                   p1.pattern, 4);
        q.amplitude = p2.amplitude;
        q.offset = p2.offset;
        q.slack = p2.slack;
        q.ok = p2.ok && p2.slack < QUICK_SLACK && fabs(p2.offset) < 1.0;
    }
    return q;
}

}  // namespace lm
}  // namespace thr
