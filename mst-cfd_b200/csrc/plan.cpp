// plan.cpp -- see plan.h
#include "plan.h"

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <numeric>
#include <parallel/algorithm>  // libstdc++ parallel mode: multi-threaded sort of the 50-100 M keys

namespace mst {

namespace {

inline uint64_t spread3(uint64_t x) {  // 21 bits -> every third bit
    x &= 0x1fffffULL;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}
inline uint64_t spread2(uint64_t x) {  // 31 bits -> every second bit
    x &= 0x7fffffffULL;
    x = (x | x << 16) & 0x0000ffff0000ffffULL;
    x = (x | x << 8) & 0x00ff00ff00ff00ffULL;
    x = (x | x << 4) & 0x0f0f0f0f0f0f0f0fULL;
    x = (x | x << 2) & 0x3333333333333333ULL;
    x = (x | x << 1) & 0x5555555555555555ULL;
    return x;
}

// Hilbert index (Skilling, "Programming the Hilbert curve", 2004): axes ->
// transposed index, in place.  b bits per axis, n axes.
inline void axes_to_transpose(uint32_t* X, int b, int n) {
    const uint32_t M = 1u << (b - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
        for (int i = 0; i < n; i++) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    for (int i = 1; i < n; i++) X[i] ^= X[i - 1];
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1)
        if (X[n - 1] & Q) t ^= Q - 1;
    for (int i = 0; i < n; i++) X[i] ^= t;
}

}  // namespace

CurveFrame bbox_frame(const mstgpu_mesh& m, int n) {
    const int D = m.dim;
    CurveFrame fr;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int c = 0; c < n; c++)
        for (int d = 0; d < D; d++) {
            double x = m.cc[(size_t)c * D + d];
            if (x == x) { lo[d] = std::min(lo[d], x); hi[d] = std::max(hi[d], x); }
        }
    // one scale for all axes: the curve's cells stay cubes on a stretched domain
    for (int d = 0; d < D; d++) { fr.lo[d] = lo[d]; fr.ext = std::max(fr.ext, hi[d] - lo[d]); }
    fr.set = true;
    return fr;
}

void curve_order(const mstgpu_mesh& m, int renumber, int n, std::vector<int32_t>& new2old, const CurveFrame* frame) {
    const int D = m.dim;
    new2old.resize(n);
    const CurveFrame fr = (frame && frame->set) ? *frame : bbox_frame(m, n);
    const double* lo = fr.lo;
    const double bits = (D == 3) ? 2097151.0 : 2147483647.0;
    double ext = fr.ext;
    if (const char* v = getenv("MSTGPU_CURVE_SCALE")) ext *= std::max(1.0, atof(v));  // experiment: lattice of the curve against the mesh
    const double sc = ext > 0 ? bits / ext : 0.0;
    std::vector<std::pair<uint64_t, int32_t>> key(n);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < n; c++) {
        uint64_t k = 0;
        uint32_t qq[3] = {0, 0, 0};
        for (int d = 0; d < D; d++) {
            double x = m.cc[(size_t)c * D + d];
            double t = (x == x) ? (x - lo[d]) * sc : 0.0;
            uint64_t q = (uint64_t)std::min(std::max(t, 0.0), bits);
            if (renumber == 2) qq[d] = (uint32_t)q;
            else k |= (D == 3 ? spread3(q) : spread2(q)) << d;
        }
        if (renumber == 2) {
            // Hilbert: consecutive cells are always neighbours -> any index range is one blob
            axes_to_transpose(qq, D == 3 ? 21 : 31, D);
            for (int d = 0; d < D; d++) k |= (D == 3 ? spread3(qq[d]) : spread2(qq[d])) << (D - 1 - d);
        }
        key[c] = {k, c};
    }
    __gnu_parallel::sort(key.begin(), key.end());
    for (int i = 0; i < n; i++) new2old[i] = key[i].second;
}

// One O(nf + nc) pass over a host mesh before anything indexes with its entries: every entry point that
// takes an mstgpu_mesh (create, partition, tile statistics, adjacency, permutation) goes through it, so a bad
// table is MSTGPU_ERR_ARG with a message instead of an out-of-bounds read.
std::string validate_mesh(const mstgpu_mesh& m) {
    const int D = m.dim;
    if (D != 2 && D != 3) return "dim must be 2 or 3";
    if (m.ncells <= 0 || m.nfaces <= 0) return "empty mesh";
    if (m.nint < 0 || m.nint > m.nfaces) return "nint out of range";
    if (!m.c0 || !m.c1 || !m.S || !m.fc || !m.eta || !m.flag || !m.ftype || !m.cc || !m.vol || !m.cf_ptr || !m.cf_idx)
        return "null mesh table";
    const int nc = m.ncells, nf = m.nfaces;
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int f = 0; f < nf; f++) {
        if (m.c0[f] < 0 || m.c0[f] >= nc) bad |= 1;
        if (m.c1[f] >= nc || m.c1[f] < -1) bad |= 2;
        if (m.ftype[f] == MSTGPU_BC_INTERIOR && m.c1[f] < 0) bad |= 4;
    }
    if (bad & 1) return "c0 out of range";
    if (bad & 2) return "c1 out of range";
    if (bad & 4) return "interior face without c1";
    if (m.cf_ptr[0] != 0) return "cf_ptr[0] must be 0";
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int c = 0; c < nc; c++) {
        const int64_t a = m.cf_ptr[c], b = m.cf_ptr[c + 1];
        if (b < a) { bad |= 8; continue; }
        if (b - a > 8) { bad |= 16; continue; }
    }
    if (bad & 8) return "cf_ptr must be non-decreasing";
    if (bad & 16) return "cells must have 1..8 faces";
    const int64_t ncf = m.cf_ptr[nc];
    if (ncf < 0 || ncf > 2 * (int64_t)nf) return "cf_ptr total inconsistent with the face count";
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int c = 0; c < nc; c++)
        for (int j = m.cf_ptr[c]; j < m.cf_ptr[c + 1]; j++) {
            const int f = m.cf_idx[j];
            if (f < 0 || f >= nf) { bad |= 32; continue; }
            if (m.c0[f] != c && m.c1[f] != c) bad |= 64;
        }
    if (bad & 32) return "cf_idx entry out of range";
    if (bad & 64) return "cf_idx lists a face that does not touch the cell";
    return "";
}

std::string build_plan(const mstgpu_mesh& m, const mstgpu_config& cfg, Plan& p, int n_owned) {
    {
        std::string verr = validate_mesh(m);
        if (!verr.empty()) return verr;
    }
    const int D = m.dim;
    const int nc = m.ncells, nf = m.nfaces;
    p.D = D; p.U = D + 2; p.nc = nc; p.nf = nf; p.nint = m.nint;

    int nslot = 0;
    for (int c = 0; c < nc; c++) nslot = std::max(nslot, m.cf_ptr[c + 1] - m.cf_ptr[c]);
    if (nslot <= 0 || nslot > 8) return "cells must have 1..8 faces";
    p.nslot = nslot;

    // ---- cell order: space-filling curve through the cell centres ------------------
    p.cell_new2old.resize(nc);
    std::iota(p.cell_new2old.begin(), p.cell_new2old.end(), 0);
    if (cfg.renumber != 0) {
        std::vector<int32_t> ord;
        curve_order(m, cfg.renumber, n_owned >= 0 ? n_owned : nc, ord, &p.frame);
        std::copy(ord.begin(), ord.end(), p.cell_new2old.begin());
    }
    p.cell_old2new.resize(nc);
    for (int i = 0; i < nc; i++) p.cell_old2new[p.cell_new2old[i]] = i;

    // ---- face order: interior by (lower new cell, higher new cell), then boundary
    {
        std::vector<std::pair<uint64_t, int32_t>> key(nf);
#pragma omp parallel for schedule(static)
        for (int f = 0; f < nf; f++) {
            uint64_t a = (uint64_t)p.cell_old2new[m.c0[f]];
            uint64_t k;
            if (m.c1[f] >= 0 && m.ftype[f] == MSTGPU_BC_INTERIOR) {
                uint64_t b = (uint64_t)p.cell_old2new[m.c1[f]];
                k = (std::min(a, b) << 31) | std::max(a, b);
            } else {
                k = (1ULL << 62) | (a << 8) | (uint64_t)(m.ftype[f] & 0xff);
            }
            key[f] = {k, f};
        }
        if (cfg.renumber != 0) __gnu_parallel::sort(key.begin(), key.end());
        else std::stable_sort(key.begin(), key.end(),
                              [](const auto& x, const auto& y) { return (x.first >> 62) < (y.first >> 62); });
        p.face_new2old.resize(nf);
        p.face_old2new.resize(nf);
        for (int i = 0; i < nf; i++) {
            p.face_new2old[i] = key[i].second;
            p.face_old2new[key[i].second] = i;
        }
    }

    // ---- per-face records -------------------------------------------------------
    int qf_from = cfg.qf_copy_from < 0 ? m.nint - 1 : cfg.qf_copy_from;
    p.fc0.resize(nf); p.fc1.resize(nf);
    p.Sd.resize((size_t)nf * D); p.dx0.resize((size_t)nf * D); p.dx1.assign((size_t)nf * D, 0.0);
    p.eta.resize(nf); p.meta.resize(nf);
    int nint_new = 0;
#pragma omp parallel for schedule(static) reduction(+ : nint_new)
    for (int i = 0; i < nf; i++) {
        const int f = p.face_new2old[i];
        const int a = m.c0[f], b = m.c1[f];
        const bool interior = (b >= 0 && m.ftype[f] == MSTGPU_BC_INTERIOR);
        p.fc0[i] = p.cell_old2new[a];
        p.fc1[i] = interior ? p.cell_old2new[b] : -1;
        uint32_t fl = 0;
        for (int d = 0; d < D; d++) {
            p.Sd[(size_t)i * D + d] = (double)m.dac[f] * m.S[(size_t)f * D + d];
            p.dx0[(size_t)i * D + d] = m.fc[(size_t)f * D + d] - m.cc[(size_t)a * D + d];
            if (interior) p.dx1[(size_t)i * D + d] = m.fc[(size_t)f * D + d] - m.cc[(size_t)b * D + d];
            if (m.flag[(size_t)f * D + d]) fl |= 1u << d;
        }
        // Qf = eta Q[c0] + (1-eta) Q[c1] below qf_from, Q[c0] from there on
        // (RhoSolver.cpp:434-440, including the off-by-one at nint-1)
        p.eta[i] = (f >= qf_from || !interior) ? 1.0 : m.eta[f];
        p.meta[i] = (uint32_t)(m.ftype[f] & 0xff) | (fl << 8);
        nint_new += interior ? 1 : 0;
    }
    (void)nint_new;

    // ---- per-cell tables --------------------------------------------------------
    p.vol.resize(nc);
    p.cf.assign((size_t)nslot * nc, -1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nc; i++) {
        const int c = p.cell_new2old[i];
        p.vol[i] = m.vol[c];
        int j = 0;
        for (int k = m.cf_ptr[c]; k < m.cf_ptr[c + 1]; k++, j++) {
            const int f = m.cf_idx[k];
            const int side = (m.c0[f] == c) ? 0 : 1;
            p.cf[(size_t)j * nc + i] = 2 * p.face_old2new[f] + side;
        }
    }

    // ---- extension tables (absent from the reference; include/mstgpu.h, mstgpu_config) -------
    if (cfg.limiter == MSTGPU_LIMITER_VENKATAKRISHNAN) {
        const double k3 = cfg.limiter_k * cfg.limiter_k * cfg.limiter_k;
        p.eps2.resize(nc);
        for (int i = 0; i < nc; i++) p.eps2[i] = k3 * (D == 3 ? p.vol[i] : p.vol[i] * std::sqrt(p.vol[i]));
    }
    if (cfg.gradient == MSTGPU_GRAD_LSQ && cfg.order == 2) {
        // Inverse-distance weighted least squares over the face neighbours,
        //   min sum_j w_j^2 (Q_j - Q_c - G.d_j)^2,  d_j = cc_j - cc_c,  w_j = 1/|d_j|;
        // a boundary face adds a mirror neighbour at d = 2 (fc - cc) with the cell's own state
        // (it enters the normal matrix only).  G = sum_j [w_j^2 M^-1 d_j] (Q_j - Q_c): the bracket
        // depends on geometry only and is computed here, once.
        p.lsq.assign((size_t)nslot * D * nc, 0.0);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < nc; i++) {
            const int c = p.cell_new2old[i];
            double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            double dj[8][3], w2[8];
            bool real[8];
            int n = 0;
            for (int k = m.cf_ptr[c]; k < m.cf_ptr[c + 1]; k++, n++) {
                const int f = m.cf_idx[k];
                const bool interior = (m.c1[f] >= 0 && m.ftype[f] == MSTGPU_BC_INTERIOR);
                const int nb = interior ? (m.c0[f] == c ? m.c1[f] : m.c0[f]) : -1;
                double d2 = 0.0;
                for (int a = 0; a < D; a++) {
                    dj[n][a] = nb >= 0 ? m.cc[(size_t)nb * D + a] - m.cc[(size_t)c * D + a]
                                       : 2.0 * (m.fc[(size_t)f * D + a] - m.cc[(size_t)c * D + a]);
                    d2 += dj[n][a] * dj[n][a];
                }
                w2[n] = 1.0 / d2;
                real[n] = nb >= 0;
                for (int a = 0; a < D; a++)
                    for (int b = 0; b < D; b++) M[a][b] += w2[n] * dj[n][a] * dj[n][b];
            }
            double inv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            if (D == 2) {
                const double det = M[0][0] * M[1][1] - M[0][1] * M[1][0];
                inv[0][0] = M[1][1] / det; inv[0][1] = -M[0][1] / det;
                inv[1][0] = -M[1][0] / det; inv[1][1] = M[0][0] / det;
            } else {
                const double c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
                const double c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
                const double c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
                const double det = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02;
                inv[0][0] = c00 / det; inv[1][0] = c01 / det; inv[2][0] = c02 / det;
                inv[0][1] = (M[0][2] * M[2][1] - M[0][1] * M[2][2]) / det;
                inv[1][1] = (M[0][0] * M[2][2] - M[0][2] * M[2][0]) / det;
                inv[2][1] = (M[0][1] * M[2][0] - M[0][0] * M[2][1]) / det;
                inv[0][2] = (M[0][1] * M[1][2] - M[0][2] * M[1][1]) / det;
                inv[1][2] = (M[0][2] * M[1][0] - M[0][0] * M[1][2]) / det;
                inv[2][2] = (M[0][0] * M[1][1] - M[0][1] * M[1][0]) / det;
            }
            for (int j = 0; j < n; j++) {
                if (!real[j]) continue;
                for (int a = 0; a < D; a++) {
                    double g = 0.0;
                    for (int b = 0; b < D; b++) g += inv[a][b] * dj[j][b];
                    p.lsq[((size_t)j * D + a) * nc + i] = w2[j] * g;
                }
            }
        }
    }
    return "";
}

}  // namespace mst
