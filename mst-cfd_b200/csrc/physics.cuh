// physics.cuh -- per-face arithmetic of the rhoSolver hot path, device side.
//
// Mirrors the semantics (not the code) of the reference's Riemann solvers
// (R = /root/reference/MST-CFD):
//   Roe    R/rhoSolver/SolverRoe.cpp:3-16 (set), :70-111 (solverAll), :114-123
//   AUSM+  R/rhoSolver/SolverAusm.cpp:3-26 (set), :53-63, :105-143
//   gas    R/work/FUNCTION.cpp:3-15
//   BCs    R/rhoSolver/RhoSolver.cpp:119-228 (1st order), :262-364 (2nd order)
//
// B200-first choices:
//  * K |L| K^-1 dU is evaluated in closed form (wave strengths), never as a
//    numerical 4x4/5x5 inverse: ~10x fewer FP64 operations, which keeps the
//    kernel under the HBM roofline instead of under the FP64 pipe.  The closed
//    form does not assume a^2 = (g-1)(H - q^2/2) (the reference's abs() can
//    break that identity), so it is the exact inverse of the reference's K.
//  * The reference calls set(L,R) once per coordinate direction with L/R
//    swapped by flagLeftRight.  Roe averages and the AUSM interface speed are
//    symmetric under the swap, so they are computed once per face; only the
//    dissipation / splitting terms see the orientation.
//  * The dimension-split flux matrix F (DIMU x DIM) is contracted with the
//    area vector in registers: phi = sum_d Sd[d] * F[:,d].  F never exists in
//    memory.
#pragma once
#include <cstdint>

namespace mst {

struct DevCfg {
    double gamma, gm1, delta, delta2, inv2delta, eor;
    double astar_fac;  // 2 (g-1)/(g+1)      SolverAusm.cpp:6
    double mu, lambda, kappa, cv, inv_cv;
    double inletQ[5];
    int32_t order, flux, viscous, limiter;  // limiter: extension (0 none, 1 Barth-Jespersen, 2 Venkatakrishnan)
};

// ---- slope limiter (extension; formulas of oracle/rho_oracle.cpp limitGradient) ------------------
// phi = min over the cell's faces of phi_j(D_j).  FP64 divisions are the expensive instruction, so the
// minimum is taken BEFORE dividing:
//  * Barth-Jespersen: min_j dmax/D_j over D_j > 0 equals dmax / max_j D_j exactly (IEEE division is
//    monotone), likewise for D_j < 0 -> two divisions per variable, bit-identical to the per-face form;
//  * Venkatakrishnan: the fractions N_j/M_j (both positive after dividing the textbook numerator and
//    denominator by D_j) are compared by cross-multiplication, one division at the end (<= 1 ulp from
//    the per-face form).
struct LimiterAcc {
    double a, b;  // BJ: largest positive / smallest negative slope;  Venkat: best numerator / denominator
    __device__ __forceinline__ void init(int mode) {
        if (mode == 1) { a = 0.0; b = 0.0; }
        else { a = 1.0; b = 1.0; }
    }
    __device__ __forceinline__ void add(int mode, double dl, double dmax, double dmin, double e2) {
        if (mode == 1) {
            a = fmax(a, dl);
            b = fmin(b, dl);
        } else if (fabs(dl) >= 1e-150) {
            const double dm = dl > 0.0 ? dmax : dmin;
            const double c = dm * dm + e2, x = dl * dm;
            const double n = c + 2.0 * x, m = c + x + 2.0 * dl * dl;
            if (n * b < a * m) { a = n; b = m; }
        }
    }
    __device__ __forceinline__ double phi(int mode, double dmax, double dmin) const {
        if (mode == 1) {
            double p = 1.0;
            if (a > 0.0) p = fmin(p, dmax / a);
            if (b < 0.0) p = fmin(p, dmin / b);
            return p;
        }
        return a / b;
    }
};

template <int D>
struct Prim {  // per-state quantities shared by all coordinate directions
    double r;      // 1/rho
    double p;      // FUNCTION.cpp:12-15
    double ht;     // FUNCTION.cpp:3-7
    double u[D];   // velocity
};

template <int D>
__device__ __forceinline__ void prim_of(const double (&q)[D + 2], const DevCfg& c, Prim<D>& s) {
    constexpr int U = D + 2;
    s.r = 1.0 / q[0];
    double m2 = q[1] * q[1];
#pragma unroll
    for (int i = 1; i < D; i++) m2 += q[i + 1] * q[i + 1];
    s.p = (q[U - 1] - 0.5 * m2 * s.r) * c.gm1;
    s.ht = (q[U - 1] + s.p) * s.r;
#pragma unroll
    for (int i = 0; i < D; i++) s.u[i] = q[i + 1] * s.r;
}

template <int D>
__device__ __forceinline__ void prim_with_r(const double (&q)[D + 2], double r, const DevCfg& c, Prim<D>& s) {
    constexpr int U = D + 2;
    s.r = r;
    double m2 = q[1] * q[1];
#pragma unroll
    for (int i = 1; i < D; i++) m2 += q[i + 1] * q[i + 1];
    s.p = (q[U - 1] - 0.5 * m2 * s.r) * c.gm1;
    s.ht = (q[U - 1] + s.p) * s.r;
#pragma unroll
    for (int i = 0; i < D; i++) s.u[i] = q[i + 1] * s.r;
}

// SolverRoe.cpp:114-123  (absolute delta, not scaled by a)
__device__ __forceinline__ double entropy_fix(double x, const DevCfg& c) {
    return (x > c.delta) ? x : (x * x + c.delta2) * c.inv2delta;
}

// phi = sum_d Sd[d] * RoeFlux_d(L,R),  (L,R) = flag[d] ? (A,B) : (B,A)
//
// Register budget first (the fused kernel lives or dies by resident warps): the central part
// 1/2 (F_A + F_B) . S does not depend on the orientation, so it is contracted with the area vector up
// front through the mass fluxes m_A.S, m_B.S -- after that A and B themselves are dead and only the
// wave strengths (theta, shear, dq0), the Roe averages and S stay live through the three upwind parts.
template <int D>
__device__ __forceinline__ void roe_contract(const double (&A)[D + 2], const double (&B)[D + 2],
                                             uint32_t flags, const double (&Sd)[D],
                                             const DevCfg& c, double (&phi)[D + 2]) {
    constexpr int U = D + 2;
    // FP64 divisions and square roots are the expensive instructions here
    // (~20 issue slots each), so the reciprocals are shared: 1/rhoA and 1/rhoB
    // come from one division, 1/(1+w) and 1/g from another, 1/a from rsqrt, and
    // 1/(rho+EOR) from a three-term series (relative error (EOR/rho)^3).
    const double rab = 1.0 / (A[0] * B[0]);
    const double ra = B[0] * rab, rb = A[0] * rab;
    double m2a = A[1] * A[1], m2b = B[1] * B[1], mna = Sd[0] * A[1], mnb = Sd[0] * B[1];
#pragma unroll
    for (int i = 1; i < D; i++) {
        m2a += A[i + 1] * A[i + 1];
        m2b += B[i + 1] * B[i + 1];
        mna += Sd[i] * A[i + 1];  // mass flux through the face, m . S
        mnb += Sd[i] * B[i + 1];
    }
    const double pa = (A[U - 1] - 0.5 * m2a * ra) * c.gm1;  // FUNCTION.cpp:12-15
    const double pb = (B[U - 1] - 0.5 * m2b * rb) * c.gm1;
    const double hta = (A[U - 1] + pa) * ra, htb = (B[U - 1] + pb) * rb;  // FUNCTION.cpp:3-7
    // Roe averages, SolverRoe.cpp:7-14 (symmetric under L<->R)
    const double w = sqrt(fabs(B[0] * ra));
    const double sw = 1.0 + w;
    double uh[D];
    double qn2 = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        uh[i] = A[i + 1] * ra + w * (B[i + 1] * rb);
        qn2 += uh[i] * uh[i];
    }
    const double Hn = hta + w * htb;
    const double gn = Hn * sw - 0.5 * qn2;  // = g (1+w)^2
    const double z = 1.0 / (sw * gn);
    const double iw = z * gn;            // 1/(1+w)
    const double ig = sw * sw * sw * z;  // 1/g
#pragma unroll
    for (int i = 0; i < D; i++) uh[i] *= iw;
    const double q2 = qn2 * iw * iw;
    const double H = Hn * iw;
    const double ya = fabs(c.gm1 * (H - 0.5 * q2));
    const double ia = rsqrt(ya);
    const double ah = ya * ia;
    // (rho + EOR) denominators of the physical fluxes, SolverRoe.cpp:87-94:
    // 1/(rho+e) = r (1 - t + t^2 - ...), t = e r
    const double ta = c.eor * ra, tb = c.eor * rb;
    const double fa = mna * (ra * (1.0 - ta + ta * ta));
    const double fb = mnb * (rb * (1.0 - tb + tb * tb));
    // central part: sum_d Sd[d] 1/2 (F_A,d + F_B,d)
    const double ph = 0.5 * (pa + pb);
    phi[0] = 0.5 * (mna + mnb);
#pragma unroll
    for (int i = 0; i < D; i++) phi[i + 1] = 0.5 * (A[i + 1] * fa + B[i + 1] * fb) + ph * Sd[i];
    phi[U - 1] = 0.5 * (hta * mna + htb * mnb);
    // wave strengths of K^-1 dq that do not depend on the direction
    const double dq0 = B[0] - A[0];
    double sh[D];  // shear strengths  dq[t+1] - u_t dq0
    double ud = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        const double dm = B[i + 1] - A[i + 1];
        ud += uh[i] * dm;
        sh[i] = dm - uh[i] * dq0;
    }
    const double theta = (0.5 * q2 * dq0 - ud + (B[U - 1] - A[U - 1])) * ig;
    const double hq2 = 0.5 * q2;
    // upwind parts: - sum_d Sd[d] (+-1/2) K |L_d| K^-1 dq
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double ss = ((flags >> d) & 1u) ? 0.5 * Sd[d] : -0.5 * Sd[d];  // 1/2 * orientation * area component
        const double beta = sh[d] * ia;
        const double lm = entropy_fix(fabs(uh[d] - ah), c);
        const double le = entropy_fix(fabs(uh[d]), c);
        const double lp = entropy_fix(fabs(uh[d] + ah), c);
        const double wm = lm * (0.5 * (theta - beta));
        const double we = le * (dq0 - theta);
        const double wp = lp * (0.5 * (theta + beta));
        const double sum = wm + we + wp;
        const double dif = ah * (wp - wm);
        double en = H * (wm + wp) + uh[d] * dif + hq2 * we;
        phi[0] -= ss * sum;
#pragma unroll
        for (int i = 0; i < D; i++) {
            double dis = uh[i] * sum;
            if (i == d) {
                dis += dif;
            } else {
                const double ws = le * sh[i];
                dis += ws;
                en += uh[i] * ws;
            }
            phi[i + 1] -= ss * dis;
        }
        phi[U - 1] -= ss * en;
    }
}

// phi = sum_d Sd[d] * AusmPlusFlux_d(L,R), quirks of SolverAusm.cpp:116-136 kept:
// one-sided `M <= 1` tests and the 3/16 term not scaled by p.
template <int D>
__device__ __forceinline__ void ausm_contract(const double (&A)[D + 2], const double (&B)[D + 2],
                                              uint32_t flags, const double (&Sd)[D],
                                              const DevCfg& c, double (&phi)[D + 2]) {
    constexpr int U = D + 2;
    Prim<D> a, b;
    prim_of<D>(A, c, a);
    prim_of<D>(B, c, b);
    // SolverAusm.cpp:6-22: a* , |U|, a~ = a*^2 / max(a*, |U|), aFace = min
    double ma2 = 0.0, mb2 = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        ma2 += A[i + 1] * A[i + 1];
        mb2 += B[i + 1] * B[i + 1];
    }
    const double asa = sqrt(a.ht * c.astar_fac), asb = sqrt(b.ht * c.astar_fac);
    const double Ua = sqrt(ma2 * a.r * a.r), Ub = sqrt(mb2 * b.r * b.r);
    const double ata = asa * asa / fmax(asa, Ua);
    const double atb = asb * asb / fmax(asb, Ub);
    const double iaf = 1.0 / fmin(ata, atb);
    // true sound speeds replace a~ in the flux, SolverAusm.cpp:112-113
    const double ca = sqrt(c.gamma * a.p * a.r), cb = sqrt(c.gamma * b.p * b.r);
    double ysum[U], ydif[U];  // aL*PhiL + aR*PhiR (symmetric), aB*PhiB - aA*PhiA
#pragma unroll
    for (int k = 0; k < U; k++) {
        double ya = ca * (k == U - 1 ? A[k] + a.p : A[k]);
        double yb = cb * (k == U - 1 ? B[k] + b.p : B[k]);
        ysum[k] = ya + yb;
        ydif[k] = yb - ya;
    }
#pragma unroll
    for (int k = 0; k < U; k++) phi[k] = 0.0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const bool fl = (flags >> d) & 1u;
        const double ML = (fl ? a.u[d] : b.u[d]) * iaf;
        const double MR = (fl ? b.u[d] : a.u[d]) * iaf;
        const double pL = fl ? a.p : b.p;
        const double pR = fl ? b.p : a.p;
        double Mp, Mm, Pp, Pm;
        {
            const double t = ML * ML - 1.0, s = ML + 1.0;
            if (ML <= 1.0) {
                Mp = 0.25 * s * s + 0.125 * t * t;
                Pp = pL * 0.25 * s * s * (2.0 - ML) + 0.1875 * ML * t * t;
            } else {
                Mp = 0.5 * (ML + fabs(ML));
                Pp = pL * 0.5 * (ML + fabs(ML)) / ML;
            }
        }
        {
            const double t = MR * MR - 1.0, s = MR - 1.0;
            if (MR <= 1.0) {
                Mm = -0.25 * s * s - 0.125 * t * t;
                Pm = pR * 0.25 * s * s * (2.0 + MR) - 0.1875 * MR * t * t;
            } else {
                Mm = 0.5 * (MR - fabs(MR));
                Pm = pR * 0.5 * (MR - fabs(MR)) / MR;
            }
        }
        const double Mf = Mm + Mp;
        const double pf = Pm + Pp;
        const double aM = fl ? fabs(Mf) : -fabs(Mf);  // |Mf| * orientation of (R-L)
#pragma unroll
        for (int k = 0; k < U; k++) {
            double F = 0.5 * (Mf * ysum[k] - aM * ydif[k]);
            if (k == d + 1) F += pf;
            phi[k] += Sd[d] * F;
        }
    }
}

template <int D>
__device__ __forceinline__ void riemann_contract(int flux, const double (&A)[D + 2],
                                                 const double (&B)[D + 2], uint32_t flags,
                                                 const double (&Sd)[D], const DevCfg& c,
                                                 double (&phi)[D + 2]) {
    if (flux == 0) roe_contract<D>(A, B, flags, Sd, c, phi);
    else ausm_contract<D>(A, B, flags, Sd, c, phi);
}

// momentum of `out` <- mirror of momentum of `in` about the unit face normal:
// m - 2 n (n.m)   (RhoSolver.cpp:150-158, 292-300, 328-336)
template <int D>
__device__ __forceinline__ void mirror_momentum(const double (&S)[D], const double (&in)[D + 2],
                                                double (&out)[D + 2]) {
    double nn = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) nn += S[i] * S[i];
    const double inn = 1.0 / sqrt(nn);
    double n[D], dot = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        n[i] = S[i] * inn;
        dot += n[i] * in[i + 1];
    }
#pragma unroll
    for (int i = 0; i < D; i++) out[i + 1] = in[i + 1] - 2.0 * n[i] * dot;
}

// Boundary ghost state.  On entry A = interior state as the zone type wants it
// (see callers), rec = reconstructed interior state, qc = cell value.
//   inlet 10    : (A = rec, B = inletQ)
//   wall 3      : (A = rec, B = A with mirrored / negated momentum)
//   symmetry 7  : (A = qc , B = rec with momentum <- mirror of qc's momentum)
//   outlet 5    : (A = B = qc)  -> pure physical flux (1st-order rule; the
//                 reference's 2nd-order outlet is undefined behaviour)
// returns false for zone types the reference has no case for (flux stays 0).
template <int D>
__device__ __forceinline__ bool boundary_states(int type, const double (&qc)[D + 2],
                                                const double (&rec)[D + 2], const double (&S)[D],
                                                const DevCfg& c, double (&A)[D + 2],
                                                double (&B)[D + 2]) {
    constexpr int U = D + 2;
    switch (type) {
        case 10:
#pragma unroll
            for (int k = 0; k < U; k++) { A[k] = rec[k]; B[k] = c.inletQ[k]; }
            return true;
        case 3:
#pragma unroll
            for (int k = 0; k < U; k++) { A[k] = rec[k]; B[k] = rec[k]; }
            if (c.viscous == 0) mirror_momentum<D>(S, A, B);
            else {
#pragma unroll
                for (int i = 0; i < D; i++) B[i + 1] = -A[i + 1];
            }
            return true;
        case 7:
#pragma unroll
            for (int k = 0; k < U; k++) { A[k] = qc[k]; B[k] = rec[k]; }
            mirror_momentum<D>(S, A, B);
            return true;
        case 5:
#pragma unroll
            for (int k = 0; k < U; k++) { A[k] = qc[k]; B[k] = qc[k]; }
            return true;
        default:
            return false;
    }
}

}  // namespace mst
