// ORACLE / TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU restatement of MST-CFD's density-based `rhoSolver` hot path, written
// from the reference's arithmetic (R = /root/reference/MST-CFD).  It is the
// checker for the CUDA path: only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library.
// The product (mst-cfd_b200/csrc) never links or calls it.
//
// Parity status: the reference cannot be compiled as-is in this image (Eigen
// is absent).  This restatement is pinned in two ways:
//   (1) against the reference's OWN sources compiled with a minimal Eigen
//       stand-in (oracle/refbuild/, outputs in oracle/_ref/, golden vectors in
//       tests/golden/ref_*.npz), and
//   (2) against external known answers (exact Sod solution, free-stream
//       preservation, conservation) in tests/test_oracle_*.py.
// 3-D, the viscous term and the 2nd-order outlet have no working reference
//  -> "parity unpinned", extension stated below.
//
// Every function cites the reference lines it follows.  Arithmetic is kept in
// the reference's evaluation order; compile with -ffp-contract=off.
//
// Build: see oracle/Makefile (g++ -O3 -fopenmp -ffp-contract=off -shared).

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

struct om_mesh {
    int32_t dim, ncells, nfaces, nint;
    const int32_t* c0;      // nfaces
    const int32_t* c1;      // nfaces, -1 on boundary faces
    const double* S;        // nfaces*dim  face area vector as stored (Face::getDirect)
    const int8_t* dac;      // nfaces      directAndCells (+1/-1)
    const double* fc;       // nfaces*dim  face centre
    const double* eta;      // nfaces      eta0
    const uint8_t* flag;    // nfaces*dim  flagLeftRight
    const int32_t* ftype;   // nfaces      zone type (2,3,5,7,10,...)
    const double* cc;       // ncells*dim  cell centre
    const double* vol;      // ncells
    const int32_t* cf_ptr;  // ncells+1    CSR cell->faces, file order
    const int32_t* cf_idx;
};

struct om_cfg {
    int32_t order;         // 1 or 2            (ACCURACY, R/include/CONST.h:6)
    int32_t flux;          // 0 Roe, 1 AUSM+    (RHOSOLVER, CONST.h:10)
    int32_t viscous;       // FLAGVISCID        (CONST.h:14): wall ghost = -momentum, + laminar term (extension)
    int32_t qf_copy_from;  // first face whose Qf = Q[c0]; reference: nint-1 (RhoSolver.cpp:438)
    int32_t nthreads;      // NUM_CPU_THREADS   (CONST.h:7), <=0: all
    int32_t pad_;
    double gamma;          // GAMMA 1.4
    double delta;          // entropyError 0.125 (SolverRoe.cpp:115)
    double eor;            // EOR 1e-10
    double mu, kappa, cv;  // VISCIDMU, TEMPK, CV (CONST.h:38-48)
    double inletQ[5];      // RhoSolver.cpp:123,266
    // ---- build-defined EXTENSION (absent from the reference): "parity unpinned" ----
    int32_t gradient;      // 0 Green-Gauss (the reference, RhoSolver.cpp:430-452), 1 weighted least squares
    int32_t limiter;       // 0 none (the reference), 1 Barth-Jespersen, 2 Venkatakrishnan
    double limiter_k;      // Venkatakrishnan K: eps^2 = (K h)^3, h = V^(1/D)
};

}  // extern "C"

namespace {

template <int D>
struct Gas {
    static constexpr int U = D + 2;
    // R/work/FUNCTION.cpp:12-15
    static inline double getP(const double* q, double gamma) {
        double m2 = q[1] * q[1] + q[2] * q[2];
        if (D == 3) m2 = m2 + q[3] * q[3];  // extension
        return (q[U - 1] - 0.5 * m2 / q[0]) * (gamma - 1);
    }
    // R/work/FUNCTION.cpp:3-7
    static inline double getht(const double* q, double gamma) {
        double p = getP(q, gamma);
        return (q[U - 1] + p) / q[0];
    }
    // R/work/FUNCTION.cpp:8-11 (2-D); extension includes w
    static inline double getT(const double* q, double cv) {
        double m2 = q[1] * q[1] + q[2] * q[2];
        if (D == 3) m2 = m2 + q[3] * q[3];
        return (q[U - 1] - 0.5 * m2 / q[0]) / q[0] / cv;
    }
};

// ---- small dense helpers (stand-in for Eigen fixed-size algebra) -----------
// 4x4: cofactor / adjugate inverse, the algorithm class Eigen uses for fixed
// 4x4 (Eigen/src/LU/InverseImpl.h, compute_inverse<.,.,4>).  5x5: Gauss-Jordan
// with partial pivoting (Eigen uses PartialPivLU above 4x4).
static void inv4(const double* m, double* inv) {
    // m row-major 4x4
    double a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3];
    double a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    double a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11];
    double a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    double b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10;
    double b02 = a00 * a13 - a03 * a10, b03 = a01 * a12 - a02 * a11;
    double b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
    double b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30;
    double b08 = a20 * a33 - a23 * a30, b09 = a21 * a32 - a22 * a31;
    double b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    double det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
    double id = 1.0 / det;
    inv[0] = (a11 * b11 - a12 * b10 + a13 * b09) * id;
    inv[1] = (-a01 * b11 + a02 * b10 - a03 * b09) * id;
    inv[2] = (a31 * b05 - a32 * b04 + a33 * b03) * id;
    inv[3] = (-a21 * b05 + a22 * b04 - a23 * b03) * id;
    inv[4] = (-a10 * b11 + a12 * b08 - a13 * b07) * id;
    inv[5] = (a00 * b11 - a02 * b08 + a03 * b07) * id;
    inv[6] = (-a30 * b05 + a32 * b02 - a33 * b01) * id;
    inv[7] = (a20 * b05 - a22 * b02 + a23 * b01) * id;
    inv[8] = (a10 * b10 - a11 * b08 + a13 * b06) * id;
    inv[9] = (-a00 * b10 + a01 * b08 - a03 * b06) * id;
    inv[10] = (a30 * b04 - a31 * b02 + a33 * b00) * id;
    inv[11] = (-a20 * b04 + a21 * b02 - a23 * b00) * id;
    inv[12] = (-a10 * b09 + a11 * b07 - a12 * b06) * id;
    inv[13] = (a00 * b09 - a01 * b07 + a02 * b06) * id;
    inv[14] = (-a30 * b03 + a31 * b01 - a32 * b00) * id;
    inv[15] = (a20 * b03 - a21 * b01 + a22 * b00) * id;
}

template <int N>
static void inv_gj(const double* m, double* inv) {
    double a[N][2 * N];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            a[i][j] = m[i * N + j];
            a[i][N + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < N; c++) {
        int piv = c;
        for (int r = c + 1; r < N; r++)
            if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (piv != c)
            for (int j = 0; j < 2 * N; j++) std::swap(a[c][j], a[piv][j]);
        double ip = 1.0 / a[c][c];
        for (int j = 0; j < 2 * N; j++) a[c][j] *= ip;
        for (int r = 0; r < N; r++) {
            if (r == c) continue;
            double f = a[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * N; j++) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) inv[i * N + j] = a[i][N + j];
}

template <int N>
static inline void inverse(const double* m, double* inv) {
    if (N == 4)
        inv4(m, inv);
    else
        inv_gj<N>(m, inv);
}

// ---- Roe (R/rhoSolver/SolverRoe.cpp) ---------------------------------------
template <int D>
struct Roe {
    static constexpr int U = D + 2;
    double L[U], R[U];
    double Dw, roeU[D], roe_ht, roe_a;
    const om_cfg* cfg;

    // SolverRoe.cpp:114-123
    inline double entropyRefine(double x) const {
        double e = cfg->delta;
        if (x > e) return x;
        return (x * x + e * e) / 2 / e;
    }
    // SolverRoe.cpp:3-16
    inline void set(const double* l, const double* r) {
        for (int k = 0; k < U; k++) { L[k] = l[k]; R[k] = r[k]; }
        const double g = cfg->gamma;
        Dw = std::sqrt(std::fabs(R[0] / L[0]));
        for (int i = 0; i < D; i++)
            roeU[i] = (L[i + 1] / L[0] + Dw * R[i + 1] / R[0]) / (1. + Dw);
        roe_ht = (Gas<D>::getht(L, g) + Dw * Gas<D>::getht(R, g)) / (1. + Dw);
        double n2 = roeU[0] * roeU[0];
        for (int i = 1; i < D; i++) n2 = n2 + roeU[i] * roeU[i];
        double nrm = std::sqrt(n2);
        roe_a = std::sqrt(std::fabs((g - 1.) * (roe_ht - 0.5 * (nrm * nrm))));
    }
    // physical flux in direction d with the (rho + EOR) denominators,
    // SolverRoe.cpp:87-88 (x), :93-94 (y); 3-D by the same pattern (extension)
    inline void physFlux(const double* q, int d, double* F) const {
        const double g = cfg->gamma, e = cfg->eor;
        double md = q[d + 1];
        F[0] = md;
        for (int i = 0; i < D; i++) {
            // 2-D literal: x: L1*L1/(L0+e)+p, L1*L2/(L0+e); y: L1*L2/(L0+e), L2*L2/(L0+e)+p
            double a = (i <= d) ? q[i + 1] : q[d + 1];
            double b = (i <= d) ? q[d + 1] : q[i + 1];
            double t = a * b / (q[0] + e);
            if (i == d) t = t + Gas<D>::getP(q, g);
            F[i + 1] = t;
        }
        F[U - 1] = Gas<D>::getht(q, g) * md;
    }
    // SolverRoe.cpp:70-111
    inline void solverAll(int d, double* out) const {
        double lam[U];
        lam[0] = entropyRefine(std::fabs(roeU[d] - roe_a));
        for (int i = 0; i < D; i++) lam[i + 1] = entropyRefine(std::fabs(roeU[d]));
        lam[D + 1] = entropyRefine(std::fabs(roeU[d] + roe_a));
        double K[U * U];
        for (int i = 0; i < U * U; i++) K[i] = 0.0;
        // columns: 0 = u-a wave, 1 = entropy, 2.. = shear (one per tangential
        // direction, ascending), U-1 = u+a wave.  2-D rows are literally
        // SolverRoe.cpp:86 (x) and :92 (y); 3-D is the extension.
        double q2 = roeU[0] * roeU[0];
        for (int i = 1; i < D; i++) q2 = q2 + roeU[i] * roeU[i];
        K[0 * U + 0] = 1; K[0 * U + 1] = 1; K[0 * U + (U - 1)] = 1;
        for (int i = 0; i < D; i++) {
            K[(i + 1) * U + 0] = (i == d) ? roeU[i] - roe_a : roeU[i];
            K[(i + 1) * U + 1] = roeU[i];
            K[(i + 1) * U + (U - 1)] = (i == d) ? roeU[i] + roe_a : roeU[i];
        }
        K[(U - 1) * U + 0] = roe_ht - roeU[d] * roe_a;
        K[(U - 1) * U + 1] = 0.5 * q2;
        K[(U - 1) * U + (U - 1)] = roe_ht + roeU[d] * roe_a;
        int col = 2;
        for (int t = 0; t < D; t++) {
            if (t == d) continue;
            K[(t + 1) * U + col] = 1;
            K[(U - 1) * U + col] = roeU[t];
            col++;
        }
        double Ki[U * U];
        inverse<U>(K, Ki);
        // roe_absA = (K * diag) * Kinv   (SolverRoe.cpp:104)
        double KL[U * U], A[U * U];
        for (int i = 0; i < U; i++)
            for (int j = 0; j < U; j++) KL[i * U + j] = K[i * U + j] * lam[j];
        for (int i = 0; i < U; i++)
            for (int j = 0; j < U; j++) {
                double s = KL[i * U + 0] * Ki[0 * U + j];
                for (int k = 1; k < U; k++) s = s + KL[i * U + k] * Ki[k * U + j];
                A[i * U + j] = s;
            }
        double FL[U], FR[U], dU[U];
        physFlux(L, d, FL);
        physFlux(R, d, FR);
        for (int k = 0; k < U; k++) dU[k] = R[k] - L[k];
        for (int i = 0; i < U; i++) {
            double s = (0.5 * A[i * U + 0]) * dU[0];
            for (int k = 1; k < U; k++) s = s + (0.5 * A[i * U + k]) * dU[k];
            out[i] = 0.5 * (FL[i] + FR[i]) - s;
        }
    }
};

// ---- AUSM+ (R/rhoSolver/SolverAusm.cpp) -------------------------------------
template <int D>
struct Ausm {
    static constexpr int U = D + 2;
    double L[U], R[U];
    double aL, aR, aFace, pL, pR;
    const om_cfg* cfg;

    // SolverAusm.cpp:3-26
    inline void set(const double* l, const double* r) {
        for (int k = 0; k < U; k++) { L[k] = l[k]; R[k] = r[k]; }
        const double g = cfg->gamma;
        double aLs = std::sqrt(2 * Gas<D>::getht(L, g) * (g - 1) / (g + 1));
        double aRs = std::sqrt(2 * Gas<D>::getht(R, g) * (g - 1) / (g + 1));
        double mL = L[1] * L[1] + L[2] * L[2];
        double mR = R[1] * R[1] + R[2] * R[2];
        if (D == 3) { mL = mL + L[3] * L[3]; mR = mR + R[3] * R[3]; }
        double UL = std::sqrt(mL / L[0] / L[0]);
        double UR = std::sqrt(mR / R[0] / R[0]);
        aL = aLs * aLs / std::max(aLs, UL);
        aR = aRs * aRs / std::max(aRs, UR);
        aFace = std::min(aL, aR);
        pL = Gas<D>::getP(L, g);
        pR = Gas<D>::getP(R, g);
    }
    // SolverAusm.cpp:53-63 + :105-143
    inline void solverAll(int d, double* out) {
        const double g = cfg->gamma;
        double machL = L[d + 1] / L[0] / aFace;
        double machR = R[d + 1] / R[0] / aFace;
        double FcaL[U], FcaR[U];
        for (int k = 0; k < U; k++) { FcaL[k] = L[k]; FcaR[k] = R[k]; }
        FcaL[U - 1] += Gas<D>::getP(L, g);
        FcaR[U - 1] += Gas<D>::getP(R, g);
        aL = std::sqrt(g * Gas<D>::getP(L, g) / L[0]);  // overwrites a-tilde, :112-113
        aR = std::sqrt(g * Gas<D>::getP(R, g) / R[0]);
        double machPlus, machMinus, pPlus, pMinus;
        // `abs(machL <= 1)` == (machL <= 1): one-sided test, literal (:116,122,129,135)
        if (machL <= 1)
            machPlus = 0.25 * (machL + 1) * (machL + 1) + 0.125 * (machL * machL - 1) * (machL * machL - 1);
        else
            machPlus = 0.5 * (machL + std::fabs(machL));
        if (machR <= 1)
            machMinus = -0.25 * (machR - 1) * (machR - 1) - 0.125 * (machR * machR - 1) * (machR * machR - 1);
        else
            machMinus = 0.5 * (machR - std::fabs(machR));
        if (machL <= 1)
            pPlus = pL * 0.25 * (machL + 1) * (machL + 1) * (2 - machL) +
                    0.1875 * machL * (machL * machL - 1) * (machL * machL - 1);  // 2nd term not scaled by pL (:130)
        else
            pPlus = pL * 0.5 * (machL + std::fabs(machL)) / machL;
        if (machR <= 1)
            pMinus = pR * 0.25 * (machR - 1) * (machR - 1) * (2 + machR) -
                     0.1875 * machR * (machR * machR - 1) * (machR * machR - 1);
        else
            pMinus = pR * 0.5 * (machR - std::fabs(machR)) / machR;
        double machFace = machMinus + machPlus;
        double pFace = pMinus + pPlus;
        for (int k = 0; k < U; k++)
            out[k] = 0.5 * (machFace * (aL * FcaL[k] + aR * FcaR[k]) -
                            std::fabs(machFace) * (aR * FcaR[k] - aL * FcaL[k]));
        out[d + 1] += pFace;
    }
};

template <int D>
struct Ctx {
    static constexpr int U = D + 2;
    om_mesh m;
    om_cfg cfg;
    std::vector<double> Qf, G, F, Fv, Gp, Gpf;  // face Q, cell grad, face flux (U x D col-major), viscous
    int nthreads;

    Ctx(const om_mesh* mesh, const om_cfg* c) : m(*mesh), cfg(*c) {
        Qf.assign((size_t)m.nfaces * U, 0.0);
        G.assign((size_t)m.ncells * U * D, 0.0);
        F.assign((size_t)m.nfaces * U * D, 0.0);  // Time.cpp:46: zero-initialised
        nthreads = cfg.nthreads;
#ifdef _OPENMP
        if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
        nthreads = 1;
#endif
    }

    inline double soutSign(int c, int f) const {
        // R/mesh/MshBlock.cpp:307-318
        return (m.c0[f] == c) ? (double)m.dac[f] : -(double)m.dac[f];
    }

    // RhoSolver.cpp:430-452 (GRAD_INTERVAL == 1)
    void updateGradFlux(const double* Q) {
        const int nint = m.nint, nf = m.nfaces, nc = m.ncells;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int i = 0; i < nint; i++) {
            const double e = m.eta[i];
            const double* a = Q + (size_t)m.c0[i] * U;
            const double* b = Q + (size_t)m.c1[i] * U;
            for (int k = 0; k < U; k++) Qf[(size_t)i * U + k] = e * a[k] + (1 - e) * b[k];
        }
        int from = cfg.qf_copy_from;
        if (from < 0) from = 0;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int i = from; i < nf; i++) {
            const double* a = Q + (size_t)m.c0[i] * U;
            for (int k = 0; k < U; k++) Qf[(size_t)i * U + k] = a[k];
        }
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int c = 0; c < nc; c++) {
            double t[U * D];
            for (int i = 0; i < U * D; i++) t[i] = 0.0;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                int f = m.cf_idx[j];
                double sg = soutSign(c, f);
                for (int d = 0; d < D; d++) {
                    double s = sg * m.S[(size_t)f * D + d];
                    for (int k = 0; k < U; k++) t[k * D + d] += Qf[(size_t)f * U + k] * s;
                }
            }
            double v = m.vol[c];
            for (int i = 0; i < U * D; i++) G[(size_t)c * U * D + i] = t[i] / v;
        }
    }

    // ---- EXTENSION (no reference code; the north star names it) -------------
    // Inverse-distance weighted least-squares gradient over the face neighbours:
    //   minimise sum_j w_j^2 (Q_j - Q_c - G.d_j)^2,  d_j = cc_j - cc_c,  w_j = 1/|d_j|.
    // A boundary face contributes a mirror neighbour at d = 2 (fc - cc) carrying the cell's own
    // state (zero normal gradient), which keeps the normal matrix regular in corner cells.
    void lsqGradient(const double* Q) {
        const int nc = m.ncells;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int c = 0; c < nc; c++) {
            double M[D][D], rhs[U][D];
            for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) M[a][b] = 0.0;
            for (int k = 0; k < U; k++) for (int a = 0; a < D; a++) rhs[k][a] = 0.0;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                const int f = m.cf_idx[j];
                const int nb = (m.c0[f] == c) ? m.c1[f] : m.c0[f];
                double d[D], d2 = 0.0;
                for (int a = 0; a < D; a++) {
                    d[a] = (nb >= 0) ? m.cc[(size_t)nb * D + a] - m.cc[(size_t)c * D + a]
                                     : 2.0 * (m.fc[(size_t)f * D + a] - m.cc[(size_t)c * D + a]);
                    d2 += d[a] * d[a];
                }
                const double w2 = 1.0 / d2;
                for (int a = 0; a < D; a++) for (int b = 0; b < D; b++) M[a][b] += w2 * d[a] * d[b];
                if (nb >= 0)
                    for (int k = 0; k < U; k++) {
                        const double dq = Q[(size_t)nb * U + k] - Q[(size_t)c * U + k];
                        for (int a = 0; a < D; a++) rhs[k][a] += w2 * d[a] * dq;
                    }
            }
            // solve M g = rhs by Gaussian elimination with partial pivoting (M is SPD, tiny)
            for (int k = 0; k < U; k++) {
                double A[D][D + 1];
                for (int a = 0; a < D; a++) { for (int b = 0; b < D; b++) A[a][b] = M[a][b]; A[a][D] = rhs[k][a]; }
                for (int i = 0; i < D; i++) {
                    int piv = i;
                    for (int r = i + 1; r < D; r++) if (std::fabs(A[r][i]) > std::fabs(A[piv][i])) piv = r;
                    if (piv != i) for (int b = 0; b <= D; b++) std::swap(A[i][b], A[piv][b]);
                    for (int r = i + 1; r < D; r++) {
                        const double fct = A[r][i] / A[i][i];
                        for (int b = i; b <= D; b++) A[r][b] -= fct * A[i][b];
                    }
                }
                double g[D];
                for (int i = D - 1; i >= 0; i--) {
                    double sum = A[i][D];
                    for (int b = i + 1; b < D; b++) sum -= A[i][b] * g[b];
                    g[i] = sum / A[i][i];
                }
                for (int a = 0; a < D; a++) G[((size_t)c * U + k) * D + a] = g[a];
            }
        }
    }

    // Slope limiter on the conserved variables, evaluated at the face centres of the cell;
    // min / max over the cell and its face neighbours.  Scales G in place: G[k][:] *= phi[k].
    //   1  Barth-Jespersen:   phi_j = min(1, dmax/D) (D > 0), min(1, dmin/D) (D < 0), 1 (D == 0)
    //   2  Venkatakrishnan:   phi_j = [(Dm^2 + e2) D + 2 D^2 Dm] / [D (Dm^2 + 2 D^2 + Dm D + e2)],
    //                         e2 = K^3 h^3, h^3 = V (3-D), V sqrt(V) (2-D); |D| < 1e-150 -> 1
    // with D = G.(fc_j - cc), Dm = dmax if D > 0 else dmin.
    void limitGradient(const double* Q) {
        const int nc = m.ncells;
        const double k3 = cfg.limiter_k * cfg.limiter_k * cfg.limiter_k;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int c = 0; c < nc; c++) {
            double qmin[U], qmax[U];
            for (int k = 0; k < U; k++) qmin[k] = qmax[k] = Q[(size_t)c * U + k];
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                const int f = m.cf_idx[j];
                const int nb = (m.c0[f] == c) ? m.c1[f] : m.c0[f];
                if (nb < 0) continue;
                for (int k = 0; k < U; k++) {
                    const double q = Q[(size_t)nb * U + k];
                    if (q < qmin[k]) qmin[k] = q;
                    if (q > qmax[k]) qmax[k] = q;
                }
            }
            const double V = m.vol[c];
            const double e2 = k3 * (D == 3 ? V : V * std::sqrt(V));
            for (int k = 0; k < U; k++) {
                const double qc = Q[(size_t)c * U + k];
                const double dmax = qmax[k] - qc, dmin = qmin[k] - qc;
                double* g = &G[((size_t)c * U + k) * D];
                double phi = 1.0;
                for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                    const int f = m.cf_idx[j];
                    double dl = 0.0;
                    for (int a = 0; a < D; a++) dl += g[a] * (m.fc[(size_t)f * D + a] - m.cc[(size_t)c * D + a]);
                    double pj = 1.0;
                    if (cfg.limiter == 1) {
                        if (dl > 0.0) pj = std::fmin(1.0, dmax / dl);
                        else if (dl < 0.0) pj = std::fmin(1.0, dmin / dl);
                    } else {
                        if (std::fabs(dl) >= 1e-150) {
                            const double dm = dl > 0.0 ? dmax : dmin;
                            const double num = (dm * dm + e2) * dl + 2.0 * dl * dl * dm;
                            const double den = dl * (dm * dm + 2.0 * dl * dl + dm * dl + e2);
                            pj = num / den;
                        }
                    }
                    if (pj < phi) phi = pj;
                }
                for (int a = 0; a < D; a++) g[a] *= phi;
            }
        }
    }

    // Largest stable explicit step of the cell-centred scheme (EXTENSION: the reference's DT is
    // fixed, Time.cpp:62):  dt = CFL * min_c V_c / sum_{f in c} (|u_c . S_f| + a_c |S_f|),
    // a = sqrt(gamma p / rho); cells whose value is not a positive finite number are skipped.
    double cflDt(const double* Q, double cfl) const {
        const int nc = m.ncells;
        double best = HUGE_VAL;
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(min : best)
        for (int c = 0; c < nc; c++) {
            const double* q = Q + (size_t)c * U;
            const double r = 1.0 / q[0];
            double m2 = 0.0;
            for (int a = 0; a < D; a++) m2 += q[a + 1] * q[a + 1];
            const double p = (q[U - 1] - 0.5 * m2 * r) * (cfg.gamma - 1.0);
            const double a = std::sqrt(cfg.gamma * p * r);
            double lam = 0.0;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                const int f = m.cf_idx[j];
                double un = 0.0, s2 = 0.0;
                for (int b = 0; b < D; b++) { un += q[b + 1] * r * m.S[(size_t)f * D + b]; s2 += m.S[(size_t)f * D + b] * m.S[(size_t)f * D + b]; }
                lam += std::fabs(un) + a * std::sqrt(s2);
            }
            const double t = m.vol[c] / lam;
            if (t > 0.0 && t < best) best = t;  // NaN and non-positive values never win
        }
        return cfl * best;
    }

    // Q[c] + G[c] * (fc - cc[c])   (RhoSolver.cpp:250)
    inline void rec(const double* Q, int c, int f, double* out) const {
        double dx[D];
        for (int d = 0; d < D; d++) dx[d] = m.fc[(size_t)f * D + d] - m.cc[(size_t)c * D + d];
        const double* g = &G[(size_t)c * U * D];
        for (int k = 0; k < U; k++) {
            double s = g[k * D + 0] * dx[0];
            for (int d = 1; d < D; d++) s = s + g[k * D + d] * dx[d];
            out[k] = Q[(size_t)c * U + k] + s;
        }
    }

    // m - 2 n (n.m), n = S/|S|   (RhoSolver.cpp:150-158, 292-300)
    inline void mirror(int f, const double* in, double* out) const {
        double n[D], nn = 0;
        for (int d = 0; d < D; d++) nn = nn + m.S[(size_t)f * D + d] * m.S[(size_t)f * D + d];
        nn = std::sqrt(nn);
        for (int d = 0; d < D; d++) n[d] = m.S[(size_t)f * D + d] / nn;
        double dot = n[0] * in[1];
        for (int d = 1; d < D; d++) dot = dot + n[d] * in[d + 1];
        for (int d = 0; d < D; d++) out[d + 1] = in[d + 1] - (2 * n[d]) * dot;
    }

    template <class SOLVER>
    inline void faceFlux(int f, const double* A, const double* B, bool useFlag) {
        // for each coordinate direction: set(L,R); F.col(d) = solverAll(d)
        SOLVER s;
        s.cfg = &cfg;
        for (int d = 0; d < D; d++) {
            bool fl = useFlag ? (m.flag[(size_t)f * D + d] != 0) : false;
            if (fl) s.set(A, B); else s.set(B, A);
            s.solverAll(d, &F[((size_t)f * D + d) * U]);
        }
    }

    // RhoSolver.cpp:90-233 (1st order) and :234-369 (2nd order)
    template <class SOLVER>
    void updateFaceFlux(const double* Q) {
        const int nf = m.nfaces;
        const bool second = (cfg.order == 2);
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int f = 0; f < nf; f++) {
            const int c = m.c0[f];
            double A[U], B[U];
            switch (m.ftype[f]) {
                case 2: {  // interior  (:99-118 / :243-261)
                    const int cb = m.c1[f];
                    if (second) { rec(Q, c, f, A); rec(Q, cb, f, B); }
                    else for (int k = 0; k < U; k++) { A[k] = Q[(size_t)c * U + k]; B[k] = Q[(size_t)cb * U + k]; }
                    faceFlux<SOLVER>(f, A, B, true);
                    break;
                }
                case 10: {  // inlet  (:119-139 / :262-282)
                    if (second) rec(Q, c, f, A);
                    else for (int k = 0; k < U; k++) A[k] = Q[(size_t)c * U + k];
                    for (int k = 0; k < U; k++) B[k] = cfg.inletQ[k];
                    faceFlux<SOLVER>(f, A, B, true);
                    break;
                }
                case 3: {  // wall  (:140-181 / :283-319)
                    if (second) rec(Q, c, f, A);
                    else for (int k = 0; k < U; k++) A[k] = Q[(size_t)c * U + k];
                    for (int k = 0; k < U; k++) B[k] = A[k];
                    if (cfg.viscous == 0) mirror(f, A, B);
                    else for (int d = 0; d < D; d++) B[d + 1] = -A[d + 1];
                    faceFlux<SOLVER>(f, A, B, true);
                    break;
                }
                case 7: {  // symmetry  (:182-213 / :320-350)
                    for (int k = 0; k < U; k++) A[k] = Q[(size_t)c * U + k];  // NOT reconstructed
                    if (second) rec(Q, c, f, B);
                    else for (int k = 0; k < U; k++) B[k] = A[k];
                    mirror(f, A, B);  // momentum of B <- mirror of A's momentum
                    faceFlux<SOLVER>(f, A, B, true);
                    break;
                }
                case 5: {  // outlet  (:214-228); 2nd order is UB in the reference -> 1st-order rule
                    for (int k = 0; k < U; k++) { A[k] = Q[(size_t)c * U + k]; B[k] = A[k]; }
                    faceFlux<SOLVER>(f, A, B, false);  // set(out,in) for every d, flag ignored
                    break;
                }
                default:
                    break;  // F keeps its previous value (zero-initialised, Time.cpp:46)
            }
        }
    }

    // ---- laminar viscous term: documented EXTENSION (reference row V is
    // broken: RhoSolver.cpp:371-429, :70-86).  Corrected formulation:
    //  Green-Gauss gradient of (u_i, T) from eta-interpolated face primitives,
    //  face gradient = eta-weighted mean, tau = mu (grad u + grad u^T) +
    //  lambda div(u) I with lambda = -0.666667 mu (CONST.h:46-47), energy flux
    //  u.tau + k grad T (CONST.h:48), Qnew += DT/V sum_f Fv . Sout.
    void viscousTerm(const double* Q, double dt, double* Qnew) {
        const int nf = m.nfaces, nc = m.ncells, nint = m.nint;
        constexpr int P = D + 1;  // primitives differentiated: u_0..u_{D-1}, T
        Gp.assign((size_t)nc * P * D, 0.0);
        Fv.assign((size_t)nf * U, 0.0);
        std::vector<double>& prim = Gpf;
        prim.assign((size_t)nf * P, 0.0);
        // face primitives from Qf (already built by updateGradFlux when order==2;
        // rebuild here so order==1 works too)
        updateQfOnly(Q);
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int f = 0; f < nf; f++) {
            const double* q = &Qf[(size_t)f * U];
            for (int d = 0; d < D; d++) prim[(size_t)f * P + d] = q[d + 1] / q[0];
            prim[(size_t)f * P + D] = Gas<D>::getT(q, cfg.cv);
        }
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int c = 0; c < nc; c++) {
            double t[P * D];
            for (int i = 0; i < P * D; i++) t[i] = 0.0;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                int f = m.cf_idx[j];
                double sg = soutSign(c, f);
                for (int d = 0; d < D; d++) {
                    double s = sg * m.S[(size_t)f * D + d];
                    for (int k = 0; k < P; k++) t[k * D + d] += prim[(size_t)f * P + k] * s;
                }
            }
            for (int i = 0; i < P * D; i++) Gp[(size_t)c * P * D + i] = t[i] / m.vol[c];
        }
        const double mu = cfg.mu, lam = -0.666667 * cfg.mu, kap = cfg.kappa;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int f = 0; f < nf; f++) {
            double gf[P * D];
            const double* ga = &Gp[(size_t)m.c0[f] * P * D];
            if (f < nint) {
                const double* gb = &Gp[(size_t)m.c1[f] * P * D];
                // same effective weight as the face values Qf (qf_copy_from folded in)
                double e = (f >= (cfg.qf_copy_from < 0 ? 0 : cfg.qf_copy_from)) ? 1.0 : m.eta[f];
                for (int i = 0; i < P * D; i++) gf[i] = e * ga[i] + (1 - e) * gb[i];
            } else {
                for (int i = 0; i < P * D; i++) gf[i] = ga[i];
            }
            double div = 0;
            for (int d = 0; d < D; d++) div = div + gf[d * D + d];
            double tau[D][D];
            for (int i = 0; i < D; i++)
                for (int j = 0; j < D; j++) {
                    tau[i][j] = mu * (gf[i * D + j] + gf[j * D + i]);
                    if (i == j) tau[i][j] = tau[i][j] + lam * div;
                }
            double* out = &Fv[(size_t)f * U];
            // contracted with the stored S (orientation applied in the gather)
            out[0] = 0;
            double en = 0;
            for (int i = 0; i < D; i++) {
                double s = 0;
                for (int j = 0; j < D; j++) s = s + tau[i][j] * m.S[(size_t)f * D + j];
                out[i + 1] = s;
            }
            for (int j = 0; j < D; j++) {
                double w = 0;
                for (int i = 0; i < D; i++) w = w + prim[(size_t)f * P + i] * tau[i][j];
                w = w + kap * gf[D * D + j];
                en = en + w * m.S[(size_t)f * D + j];
            }
            out[U - 1] = en;
        }
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int c = 0; c < nc; c++) {
            double acc[U];
            for (int k = 0; k < U; k++) acc[k] = 0;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                int f = m.cf_idx[j];
                double sg = soutSign(c, f);
                for (int k = 0; k < U; k++) acc[k] += sg * Fv[(size_t)f * U + k];
            }
            double s = dt / m.vol[c];
            for (int k = 0; k < U; k++) Qnew[(size_t)c * U + k] += s * acc[k];
        }
    }

    void updateQfOnly(const double* Q) {
        const int nint = m.nint, nf = m.nfaces;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int i = 0; i < nint; i++) {
            const double e = m.eta[i];
            const double* a = Q + (size_t)m.c0[i] * U;
            const double* b = Q + (size_t)m.c1[i] * U;
            for (int k = 0; k < U; k++) Qf[(size_t)i * U + k] = e * a[k] + (1 - e) * b[k];
        }
        int from = cfg.qf_copy_from < 0 ? 0 : cfg.qf_copy_from;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int i = from; i < nf; i++) {
            const double* a = Q + (size_t)m.c0[i] * U;
            for (int k = 0; k < U; k++) Qf[(size_t)i * U + k] = a[k];
        }
    }

    // RhoSolver.cpp:37-68
    void solve(double dt, const double* Qold, double* Qnew) {
        if (cfg.order == 2) {
            updateGradFlux(Qold);                    // the reference's gradient (also fills Qf)
            if (cfg.gradient == 1) lsqGradient(Qold);  // EXTENSION: replaces G
            if (cfg.limiter != 0) limitGradient(Qold); // EXTENSION: scales G
        }
        if (cfg.flux == 0) updateFaceFlux<Roe<D>>(Qold);
        else updateFaceFlux<Ausm<D>>(Qold);
        const int nc = m.ncells;
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int c = 0; c < nc; c++) {
            double acc[U];
            for (int k = 0; k < U; k++) acc[k] = 0.0;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                int f = m.cf_idx[j];
                double sg = soutSign(c, f);
                for (int d = 0; d < D; d++) {
                    double s = sg * m.S[(size_t)f * D + d];
                    const double* col = &F[((size_t)f * D + d) * U];
                    for (int k = 0; k < U; k++) acc[k] += s * col[k];
                }
            }
            double s = dt / m.vol[c];
            for (int k = 0; k < U; k++) Qnew[(size_t)c * U + k] = Qold[(size_t)c * U + k] - s * acc[k];
        }
        if (cfg.viscous) viscousTerm(Qold, dt, Qnew);
    }

    // ---- EXTENSION: implicit operator for the LU-SGS sweeps (the reference has
    // the solver, R/lusolver/SparseSolver.cpp:54-104, but no rhoSolver call site, so WHAT it solves is
    // build-defined).  Linearised backward Euler with first-order flux Jacobians (Yoon-Jameson):
    //   [V_i/dt I + sum_f 1/2 (A(Q_i,S) + lam_f I)] dQ_i + sum_{f interior} 1/2 (A(Q_j,S) - lam_f I) dQ_j = -R_i
    // S = outward area vector of cell i on face f, A(Q,S) = d(F(Q).S)/dQ (perfect gas), lam(Q,S) =
    // |u.S| + a|S|, lam_f = max over the two cells (own cell on boundary faces), R_i = sum_f Phi_f the
    // explicit residual of solve().  dt -> 0 recovers the explicit step.
    static void fluxJacobian(const double* q, const double* S, double gamma, double* A /* U x U row-major */) {
        const double r = 1.0 / q[0];
        double u[D], un = 0.0, q2 = 0.0;
        for (int a = 0; a < D; a++) { u[a] = q[a + 1] * r; un += u[a] * S[a]; q2 += u[a] * u[a]; }
        const double g1 = gamma - 1.0;
        const double phi = 0.5 * g1 * q2;
        const double p = (q[U - 1] - 0.5 * q[0] * q2) * g1;
        const double H = (q[U - 1] + p) * r;
        for (int i = 0; i < U * U; i++) A[i] = 0.0;
        for (int b = 0; b < D; b++) A[0 * U + 1 + b] = S[b];
        for (int a = 0; a < D; a++) {
            A[(1 + a) * U + 0] = S[a] * phi - u[a] * un;
            for (int b = 0; b < D; b++) A[(1 + a) * U + 1 + b] = u[a] * S[b] - g1 * u[b] * S[a] + (a == b ? un : 0.0);
            A[(1 + a) * U + U - 1] = g1 * S[a];
        }
        A[(U - 1) * U + 0] = (phi - H) * un;
        for (int b = 0; b < D; b++) A[(U - 1) * U + 1 + b] = H * S[b] - g1 * u[b] * un;
        A[(U - 1) * U + U - 1] = gamma * un;
    }
    static double spectralRadius(const double* q, const double* S, double gamma) {
        const double r = 1.0 / q[0];
        double un = 0.0, q2 = 0.0, s2 = 0.0;
        for (int a = 0; a < D; a++) { un += q[a + 1] * r * S[a]; q2 += q[a + 1] * q[a + 1]; s2 += S[a] * S[a]; }
        const double p = (q[U - 1] - 0.5 * q2 * r) * (gamma - 1.0);
        return std::fabs(un) + std::sqrt(gamma * p * r) * std::sqrt(s2);
    }

    // R_i = sum_f sum_d Sout[d] F[:,d]: the gather of solve() without the update
    void residualVector(const double* Q, double* R) {
        if (cfg.order == 2) {
            updateGradFlux(Q);
            if (cfg.gradient == 1) lsqGradient(Q);
            if (cfg.limiter != 0) limitGradient(Q);
        }
        if (cfg.flux == 0) updateFaceFlux<Roe<D>>(Q);
        else updateFaceFlux<Ausm<D>>(Q);
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int c = 0; c < m.ncells; c++) {
            double acc[U];
            for (int k = 0; k < U; k++) acc[k] = 0.0;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                int f = m.cf_idx[j];
                double sg = soutSign(c, f);
                for (int d = 0; d < D; d++) {
                    double s = sg * m.S[(size_t)f * D + d];
                    const double* colp = &F[((size_t)f * D + d) * U];
                    for (int k = 0; k < U; k++) acc[k] += s * colp[k];
                }
            }
            for (int k = 0; k < U; k++) R[(size_t)c * U + k] = acc[k];
        }
    }

    // CSR by row, columns ascending, blocks U x U row-major; b = -R.  rowptr [nc+1]; col / val may be
    // null for a sizing call.  Returns the number of blocks.
    int64_t implicitSystem(double dt, const double* Q, int32_t* rowptr, int32_t* col, double* val, double* b) {
        const int nc = m.ncells;
        rowptr[0] = 0;
        for (int c = 0; c < nc; c++) {
            int cnt = 1;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                const int f = m.cf_idx[j];
                if (m.ftype[f] == 2 && m.c1[f] >= 0) cnt++;
            }
            rowptr[c + 1] = rowptr[c] + cnt;
        }
        if (!col || !val || !b) return rowptr[nc];
        std::vector<double> R((size_t)nc * U);
        residualVector(Q, R.data());
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int c = 0; c < nc; c++) {
            int32_t* cols = col + rowptr[c];
            double* blk = val + (size_t)rowptr[c] * U * U;
            int n = 0;
            cols[n++] = c;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                const int f = m.cf_idx[j];
                if (m.ftype[f] == 2 && m.c1[f] >= 0) cols[n++] = (m.c0[f] == c) ? m.c1[f] : m.c0[f];
            }
            std::sort(cols, cols + n);
            for (size_t i = 0; i < (size_t)n * U * U; i++) blk[i] = 0.0;
            auto at = [&](int cc) { return blk + (size_t)(std::lower_bound(cols, cols + n, cc) - cols) * U * U; };
            double* Dg = at(c);
            for (int k = 0; k < U; k++) Dg[k * U + k] = m.vol[c] / dt;
            const double* qi = Q + (size_t)c * U;
            for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
                const int f = m.cf_idx[j];
                const double sg = soutSign(c, f);
                double S[D], A[U * U];
                for (int d = 0; d < D; d++) S[d] = sg * m.S[(size_t)f * D + d];
                const bool interior = (m.ftype[f] == 2 && m.c1[f] >= 0);
                const int nb = interior ? ((m.c0[f] == c) ? m.c1[f] : m.c0[f]) : -1;
                double lam = spectralRadius(qi, S, cfg.gamma);
                if (nb >= 0) lam = std::fmax(lam, spectralRadius(Q + (size_t)nb * U, S, cfg.gamma));
                fluxJacobian(qi, S, cfg.gamma, A);
                for (int i = 0; i < U * U; i++) Dg[i] += 0.5 * A[i];
                for (int k = 0; k < U; k++) Dg[k * U + k] += 0.5 * lam;
                if (nb >= 0) {
                    double* O = at(nb);
                    fluxJacobian(Q + (size_t)nb * U, S, cfg.gamma, A);
                    for (int i = 0; i < U * U; i++) O[i] += 0.5 * A[i];
                    for (int k = 0; k < U; k++) O[k * U + k] -= 0.5 * lam;
                }
            }
            for (int k = 0; k < U; k++) b[(size_t)c * U + k] = -R[(size_t)c * U + k];
        }
        return rowptr[nc];
    }

    // R/time/Time.cpp:69-76: signed denominator, NaN never wins, +inf can
    void residual(const double* Qold, const double* Qnew, double* r) const {
        // the reference loop is serial; max is exact under any association, so
        // the OpenMP reduction below returns the same value
        double rr[U];
        for (int k = 0; k < U; k++) rr[k] = 0.0;
#pragma omp parallel num_threads(nthreads)
        {
            double loc[U];
            for (int k = 0; k < U; k++) loc[k] = 0.0;
#pragma omp for schedule(static) nowait
            for (int c = 0; c < m.ncells; c++)
                for (int k = 0; k < U; k++) {
                    double x = std::fabs(Qnew[(size_t)c * U + k] - Qold[(size_t)c * U + k]) / Qold[(size_t)c * U + k];
                    if (loc[k] < x) loc[k] = x;
                }
#pragma omp critical
            for (int k = 0; k < U; k++)
                if (rr[k] < loc[k]) rr[k] = loc[k];
        }
        for (int k = 0; k < U; k++) r[k] = rr[k];
    }
};

struct Handle {
    int dim;
    void* p;
};

}  // namespace

extern "C" {

void* oracle_create(const om_mesh* m, const om_cfg* c) {
    Handle* h = new Handle;
    h->dim = m->dim;
    if (m->dim == 2) h->p = new Ctx<2>(m, c);
    else if (m->dim == 3) h->p = new Ctx<3>(m, c);
    else { delete h; return nullptr; }
    return h;
}

void oracle_destroy(void* hv) {
    Handle* h = (Handle*)hv;
    if (!h) return;
    if (h->dim == 2) delete (Ctx<2>*)h->p; else delete (Ctx<3>*)h->p;
    delete h;
}

int oracle_nthreads(void* hv) {
    Handle* h = (Handle*)hv;
    return h->dim == 2 ? ((Ctx<2>*)h->p)->nthreads : ((Ctx<3>*)h->p)->nthreads;
}

// one RhoSolver::solve(): Qold -> Qnew (both ncells*U, AoS like VCTDIMU[])
int oracle_solve(void* hv, double dt, const double* Qold, double* Qnew) {
    Handle* h = (Handle*)hv;
    if (h->dim == 2) ((Ctx<2>*)h->p)->solve(dt, Qold, Qnew);
    else ((Ctx<3>*)h->p)->solve(dt, Qold, Qnew);
    return 0;
}

// Time::goNextTimeStep() x nsteps: solve, residual, new->old.  Q in/out.
// resid (optional) receives nsteps*U values.
int oracle_run(void* hv, double dt, int nsteps, double* Q, double* resid) {
    Handle* h = (Handle*)hv;
    const int U = h->dim + 2;
    size_t n = (size_t)(h->dim == 2 ? ((Ctx<2>*)h->p)->m.ncells : ((Ctx<3>*)h->p)->m.ncells) * U;
    std::vector<double> Qn(n);
    for (int s = 0; s < nsteps; s++) {
        oracle_solve(hv, dt, Q, Qn.data());
        if (resid) {
            if (h->dim == 2) ((Ctx<2>*)h->p)->residual(Q, Qn.data(), resid + (size_t)s * U);
            else ((Ctx<3>*)h->p)->residual(Q, Qn.data(), resid + (size_t)s * U);
        }
        std::memcpy(Q, Qn.data(), n * sizeof(double));  // RhoSolver.cpp:513-517
    }
    return 0;
}

// EXTENSION: CFL time step of the state Q (the reference's DT is a macro, Time.cpp:62)
double oracle_cfl_dt(void* hv, double cfl, const double* Q) {
    Handle* h = (Handle*)hv;
    return h->dim == 2 ? ((Ctx<2>*)h->p)->cflDt(Q, cfl) : ((Ctx<3>*)h->p)->cflDt(Q, cfl);
}

// nsteps steps, each with dt = oracle_cfl_dt of its start state; dts (optional) receives them
int oracle_run_cfl(void* hv, double cfl, int nsteps, double* Q, double* dts) {
    Handle* h = (Handle*)hv;
    const int U = h->dim + 2;
    size_t n = (size_t)(h->dim == 2 ? ((Ctx<2>*)h->p)->m.ncells : ((Ctx<3>*)h->p)->m.ncells) * U;
    std::vector<double> Qn(n);
    for (int s = 0; s < nsteps; s++) {
        const double dt = oracle_cfl_dt(hv, cfl, Q);
        if (dts) dts[s] = dt;
        oracle_solve(hv, dt, Q, Qn.data());
        std::memcpy(Q, Qn.data(), n * sizeof(double));
    }
    return 0;
}

// EXTENSION: block system of one implicit step (see Ctx::implicitSystem); col / val / b may be NULL to size
int64_t oracle_implicit_system(void* hv, double dt, const double* Q, int32_t* rowptr, int32_t* col, double* val, double* b) {
    Handle* h = (Handle*)hv;
    return h->dim == 2 ? ((Ctx<2>*)h->p)->implicitSystem(dt, Q, rowptr, col, val, b)
                       : ((Ctx<3>*)h->p)->implicitSystem(dt, Q, rowptr, col, val, b);
}

// stage probes (valid after oracle_solve): Qf nfaces*U, G ncells*U*D
// (row-major U x D), F nfaces*D*U (Eigen column-major U x D == [d][k])
int oracle_probe(void* hv, double* Qf, double* G, double* F) {
    Handle* h = (Handle*)hv;
    if (h->dim == 2) {
        Ctx<2>* c = (Ctx<2>*)h->p;
        if (Qf) std::memcpy(Qf, c->Qf.data(), c->Qf.size() * 8);
        if (G) std::memcpy(G, c->G.data(), c->G.size() * 8);
        if (F) std::memcpy(F, c->F.data(), c->F.size() * 8);
    } else {
        Ctx<3>* c = (Ctx<3>*)h->p;
        if (Qf) std::memcpy(Qf, c->Qf.data(), c->Qf.size() * 8);
        if (G) std::memcpy(G, c->G.data(), c->G.size() * 8);
        if (F) std::memcpy(F, c->F.data(), c->F.size() * 8);
    }
    return 0;
}

// single Riemann solves, for unit tests of rows F and G
int oracle_riemann(int dim, int flux, const om_cfg* cfg, const double* L, const double* R, int d, double* out) {
    if (dim == 2) {
        if (flux == 0) { Roe<2> s; s.cfg = cfg; s.set(L, R); s.solverAll(d, out); }
        else { Ausm<2> s; s.cfg = cfg; s.set(L, R); s.solverAll(d, out); }
    } else {
        if (flux == 0) { Roe<3> s; s.cfg = cfg; s.set(L, R); s.solverAll(d, out); }
        else { Ausm<3> s; s.cfg = cfg; s.set(L, R); s.solverAll(d, out); }
    }
    return 0;
}

}  // extern "C"
