// tiles.cpp -- see tiles.h
#include "tiles.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "tile_layout.h"

namespace mst {

namespace {

// tiny open-addressing map int32 -> int32 with O(1) reset (generation stamps)
struct SmallMap {
    std::vector<int32_t> key, val;
    std::vector<uint32_t> gen;
    uint32_t cur = 1, mask;
    explicit SmallMap(int log2cap) : key(1u << log2cap), val(1u << log2cap), gen(1u << log2cap, 0), mask((1u << log2cap) - 1) {}
    void reset() { cur++; }
    static uint32_t h(int32_t k) { uint32_t x = (uint32_t)k * 2654435761u; return x ^ (x >> 15); }
    int32_t find(int32_t k) const {
        for (uint32_t i = h(k) & mask;; i = (i + 1) & mask) {
            if (gen[i] != cur) return -1;
            if (key[i] == k) return val[i];
        }
    }
    void put(int32_t k, int32_t v) {
        for (uint32_t i = h(k) & mask;; i = (i + 1) & mask)
            if (gen[i] != cur) { gen[i] = cur; key[i] = k; val[i] = v; return; }
    }
};

struct TileScratch {
    SmallMap cells{14}, faces{14};
    std::vector<int32_t> ring, flist;  // ring cell ids; local face -> device face id
};

inline int up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

std::string build_tiles(const Plan& p, int n_update, int T, int order, TilePack& tp, int ext, int fit_faces) {
    if (order != 2) ext = 0;
    const int limiter = ext & 1, visc = ext & 2;
    const int D = p.D, nc = p.nc, nslot = p.nslot, NS = nslot + 1;
    if (T < 32 || T > 2048 || (T & 1)) return "tile size must be even, 32..2048";
    if (n_update <= 0 || n_update > nc) return "n_update out of range";
    tp.T = T; tp.order = order; tp.D = D; tp.nslot = nslot; tp.ext = ext;
    const bool lsq = !p.lsq.empty();
    // tile t = cells [tb[t], tb[t+1]).  Fixed size T, or (fit_faces > 0) grown cell by cell while the tile's flux
    // faces fit `fit_faces` (a multiple of the CTA size: every warp of the CTA then makes the same number of
    // trips through phase 2 -- no trip with a handful of live lanes at the end) and its cells fit T.  Tiles
    // start at even cells (the bulk copies of the state move 16-byte units = 2 rows of 40 bytes).
    std::vector<int32_t> tb;
    if (fit_faces > 0) {
        tb.push_back(0);
        int cb = 0, nfb = 0;
        for (int c = 0; c < n_update; c++) {
            int add = 0;
            for (int j = 0; j < nslot; j++) {
                const int v = p.cf[(size_t)j * nc + c];
                if (v < 0) continue;
                const int f = v >> 1, nb = (v & 1) ? p.fc0[f] : p.fc1[f];
                if (!(nb >= cb && nb < c)) add++;  // not yet listed by an owned neighbour
            }
            if (c - cb >= 2 && (nfb + add > fit_faces || c - cb >= T) ) {
                // close the tile at the last even size that fits
                int end = c;
                if ((end - cb) & 1) end--;
                tb.push_back(end);
                // faces of the cells [end, c] that were counted for the old tile start a new count
                cb = end; nfb = 0;
                for (int cc = cb; cc < c; cc++)
                    for (int j = 0; j < nslot; j++) {
                        const int v = p.cf[(size_t)j * nc + cc];
                        if (v < 0) continue;
                        const int f = v >> 1, nb = (v & 1) ? p.fc0[f] : p.fc1[f];
                        if (!(nb >= cb && nb < cc)) nfb++;
                    }
                add = 0;
                for (int j = 0; j < nslot; j++) {
                    const int v = p.cf[(size_t)j * nc + c];
                    if (v < 0) continue;
                    const int f = v >> 1, nb = (v & 1) ? p.fc0[f] : p.fc1[f];
                    if (!(nb >= cb && nb < c)) add++;
                }
            }
            nfb += add;
        }
        tb.push_back(n_update);
    } else {
        for (int c = 0; c < n_update; c += T) tb.push_back(c);
        tb.push_back(n_update);
    }
    tp.ntiles = (int)tb.size() - 1;
    tp.desc.assign(tp.ntiles, TileDesc{});
    const int nt = tp.ntiles;

    auto nb_of = [&](int c, int j, int& f, int& side) -> int {
        const int v = p.cf[(size_t)j * nc + c];
        if (v < 0) { f = -1; side = 0; return -2; }
        f = v >> 1; side = v & 1;
        return side ? p.fc0[f] : p.fc1[f];
    };

    // pass 0: sizes; pass 1: fill.  The traversal is identical in both passes.
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int64_t ro = 0, po = 0;
            for (int t = 0; t < nt; t++) {
                TileDesc& d = tp.desc[t];
                const TileLayout L = tile_layout(D, order, nslot, d.n_own, d.n_r1, d.n_r2, d.nFB, ext);
                d.ring_off = ro; d.pk_off = po;
                ro += up(d.n_r1 + d.n_r2, 4);
                po += L.pk_bytes;
                tp.max_smem = std::max(tp.max_smem, (size_t)L.total);
                tp.sum_r1 += d.n_r1; tp.sum_r2 += d.n_r2; tp.sum_FB += d.nFB; tp.sum_FA += d.nFA;
            }
            tp.ring.assign(ro, 0);
            tp.packets.assign((size_t)po, 0);
        }
        std::string err;
#pragma omp parallel
        {
            TileScratch s;
#pragma omp for schedule(dynamic, 64)
            for (int t = 0; t < nt; t++) {
                TileDesc& d = tp.desc[t];
                const int cb = tb[t], ce = tb[t + 1], n_own = ce - cb;
                s.cells.reset(); s.faces.reset(); s.ring.clear(); s.flist.clear();
                auto owned = [&](int g) { return g >= cb && g < ce; };
                auto local_of = [&](int g) -> int { return owned(g) ? g - cb : s.cells.find(g); };
                int n_r1 = 0, n_r2 = 0;
                // ring 1: face neighbours of owned cells
                for (int c = cb; c < ce; c++)
                    for (int j = 0; j < nslot; j++) {
                        int f, side;
                        const int nb = nb_of(c, j, f, side);
                        if (nb >= 0 && !owned(nb) && s.cells.find(nb) < 0) {
                            s.cells.put(nb, n_own + n_r1++);
                            s.ring.push_back(nb);
                        }
                    }
                // ring 2: face neighbours of ring-1 cells (their reconstruction stencil), order 2 only
                if (order == 2)
                    for (int i = 0; i < n_r1; i++) {
                        const int r = s.ring[i];
                        for (int j = 0; j < nslot; j++) {
                            int f, side;
                            const int nb = nb_of(r, j, f, side);
                            if (nb >= 0 && !owned(nb) && s.cells.find(nb) < 0) {
                                s.cells.put(nb, n_own + n_r1 + n_r2++);
                                s.ring.push_back(nb);
                            }
                        }
                    }
                // flux faces: faces of owned cells, each once
                for (int c = cb; c < ce; c++)
                    for (int j = 0; j < nslot; j++) {
                        int f, side;
                        const int nb = nb_of(c, j, f, side);
                        if (f < 0) continue;
                        if (nb >= 0 && owned(nb) && side == 1) continue;  // the c0-side cell lists it
                        if (s.faces.find(f) >= 0) continue;
                        s.faces.put(f, (int)s.flist.size());
                        s.flist.push_back(f);
                    }
                const int nFB = (int)s.flist.size();
                if (pass == 0) {
                    d.cb = cb; d.n_own = n_own; d.n_r1 = n_r1; d.n_r2 = n_r2; d.nFB = nFB; d.nFA = nFB;
                    if (n_own + n_r1 + n_r2 >= 0xFFFF || nFB >= 0x7FFF || n_own + n_r1 + n_r2 > 12000 || nFB > 12000) {
#pragma omp critical
                        err = "tile too large for 16-bit local indices";
                    }
                    continue;
                }
                // ---- fill ---------------------------------------------------------
                const TileLayout L = tile_layout(D, order, nslot, n_own, n_r1, n_r2, nFB, ext);
                unsigned char* pk = tp.packets.data() + d.pk_off;
                double* w = reinterpret_cast<double*>(pk + L.w);
                uint32_t* idx = reinterpret_cast<uint32_t*>(pk + L.idx);
                double* fSd = reinterpret_cast<double*>(pk + L.fSd);
                uint32_t* fmeta = reinterpret_cast<uint32_t*>(pk + L.fmeta);
                uint16_t* slots = reinterpret_cast<uint16_t*>(pk + L.slots);
                double* cvol = reinterpret_cast<double*>(pk + L.cvol);
                const uint32_t nFBp = L.nFBp, ncp = L.ncp;
                for (size_t i = 0; i < s.ring.size(); i++) tp.ring[d.ring_off + i] = s.ring[i];
                for (uint32_t i = 0; i < (uint32_t)nslot * ncp; i++) slots[i] = 0xFFFF;
                for (uint32_t i = 0; i < ncp; i++) cvol[i] = 1.0;
                for (uint32_t i = 0; i < (order == 2 ? (uint32_t)NS - 1u : 1u) * nFBp; i++) idx[i] = 0;  // padded faces read cell 0
                for (int lc = 0; lc < n_own; lc++) {
                    const int g = cb + lc;
                    cvol[lc] = 1.0 / p.vol[g];  // the kernel multiplies: DT * (1/V)
                    for (int j = 0; j < nslot; j++) {
                        int f, side;
                        nb_of(g, j, f, side);
                        if (f < 0) continue;
                        slots[(size_t)j * ncp + lc] = (uint16_t)((s.faces.find(f) << 1) | side);
                    }
                }
                // Second-order reconstruction as a fixed stencil (RhoSolver.cpp:250, 434-452):
                //   rec(c,f) = Q_c + G_c (fc_f - cc_c),  G_c = (1/V) sum_j Qf_j (x) Sout_j,
                //   Qf_j = eta Q[c0] + (1-eta) Q[c1]  (Q_c on boundary faces)
                // is linear in the states of c and its face neighbours:
                //   rec = b0 Q_c + sum_j bj Q_nb(j),  sigma_j = Sout_j.(fc_f - cc_c)/V,
                //   b0 = 1 + sum_j wself_j sigma_j,  bj = wnb_j sigma_j.
                // The weights depend on geometry only and are computed here, once.
                // Slot 1 of a side's stencil is always the cell across the face itself (the other
                // side's own cell), so the kernel loads the two cells of the face once for both sides.
                // gradient of cell c projected on dx, as weights of the stencil (c, face neighbours in the
                // cell's own face order): G_c.dx = (beta[0] - init0) Q_c + sum_j beta[1+j] Q_nb(j)
                auto project = [&](int c, const double* dx, double init0, double* beta, int* cells) {
                    const double V = p.vol[c];
                    beta[0] = init0;
                    cells[0] = local_of(c);
                    for (int j = 0; j < nslot; j++) {
                        beta[1 + j] = 0.0;
                        cells[1 + j] = cells[0];
                        int g, side;
                        const int nb = nb_of(c, j, g, side);
                        if (g < 0) continue;
                        if (nb >= 0) cells[1 + j] = local_of(nb);
                        if (lsq) {
                            // extension: least-squares gradient, G = sum_j lsq_j (Q_nb(j) - Q_c) (plan.h)
                            if (nb < 0) continue;
                            double b = 0.0;
                            for (int k = 0; k < D; k++) b += p.lsq[((size_t)j * D + k) * nc + c] * dx[k];
                            beta[1 + j] = b;
                            beta[0] -= b;
                            continue;
                        }
                        double dot = 0.0;
                        for (int k = 0; k < D; k++) dot += p.Sd[(size_t)g * D + k] * dx[k];
                        const double sigma = (side ? -dot : dot) / V;
                        const double e = p.eta[g];
                        if (nb >= 0) {
                            beta[0] += (side ? (1.0 - e) : e) * sigma;
                            beta[1 + j] = (side ? e : (1.0 - e)) * sigma;
                        } else {
                            beta[0] += sigma;
                        }
                    }
                };
                auto stencil = [&](int c, int fself, const double* dx, double* beta, int* cells) {
                    project(c, dx, 1.0, beta, cells);
                    int jself = 0;
                    for (int j = 0; j < nslot; j++) {
                        int g, side;
                        nb_of(c, j, g, side);
                        if (g == fself) jself = j;
                    }
                    std::swap(beta[1], beta[1 + jself]);
                    std::swap(cells[1], cells[1 + jself]);
                };
                if (ext) {
                    // face-neighbour ids of the stencil cells (own + ring 1); own id where a slot has no neighbour
                    uint16_t* lid = reinterpret_cast<uint16_t*>(pk + L.lid);
                    const uint32_t nCLp = L.nCLp;
                    for (uint32_t i = 0; i < nCLp; i++)
                        for (int m = 0; m < nslot; m++) {
                            int v = 0;
                            if ((int)i < n_own + n_r1) {
                                const int c = (int)i < n_own ? cb + (int)i : s.ring[i - n_own];
                                int g, side;
                                const int nb = nb_of(c, m, g, side);
                                v = nb >= 0 ? local_of(nb) : (int)i;
                            }
                            lid[(size_t)m * nCLp + i] = (uint16_t)v;
                        }
                }
                if (visc) {
                    // viscous term: Green-Gauss gradient of the face primitives needs, per face slot of a stencil
                    // cell, the two interpolation weights of the face state and the outward area vector / V
                    double* vw = reinterpret_cast<double*>(pk + L.vw);
                    double* feta = reinterpret_cast<double*>(pk + L.feta);
                    const uint32_t nCLp = L.nCLp;
                    const int W = 2 + D;
                    for (int i = 0; i < n_own + n_r1; i++) {
                        const int c = i < n_own ? cb + i : s.ring[i - n_own];
                        for (int j = 0; j < nslot; j++) {
                            int g, side;
                            const int nb = nb_of(c, j, g, side);
                            double e0 = 1.0, e1 = 0.0, S[3] = {0.0, 0.0, 0.0};  // pad slot: contributes nothing
                            if (g >= 0) {
                                const double e = p.eta[g];
                                if (nb >= 0) { e0 = side ? (1.0 - e) : e; e1 = side ? e : (1.0 - e); }
                                for (int k = 0; k < D; k++) S[k] = (side ? -1.0 : 1.0) * p.Sd[(size_t)g * D + k] / p.vol[c];
                            }
                            vw[((size_t)j * W + 0) * nCLp + i] = e0;
                            vw[((size_t)j * W + 1) * nCLp + i] = e1;
                            for (int k = 0; k < D; k++) vw[((size_t)j * W + 2 + k) * nCLp + i] = S[k];
                        }
                    }
                    for (int i = n_own + n_r1; i < (int)nCLp; i++)
                        for (int j = 0; j < nslot; j++) vw[((size_t)j * W + 0) * nCLp + i] = 1.0;
                    for (int lf = 0; lf < nFB; lf++) feta[lf] = p.eta[s.flist[lf]];
                    for (uint32_t lf = (uint32_t)nFB; lf < L.nFBp; lf++) feta[lf] = 1.0;
                }
                if (limiter) {
                    // limiter tables of the cells whose reconstruction the tile evaluates: own + ring 1
                    double* lw = reinterpret_cast<double*>(pk + L.lw);
                    double* le2 = reinterpret_cast<double*>(pk + L.le2);
                    const uint32_t nCLp = L.nCLp;
                    for (int i = 0; i < n_own + n_r1; i++) {
                        const int c = i < n_own ? cb + i : s.ring[i - n_own];
                        le2[i] = p.eps2.empty() ? 0.0 : p.eps2[c];
                        for (int j = 0; j < nslot; j++) {
                            double beta[9];
                            int cells[9];
                            int g, side;
                            nb_of(c, j, g, side);
                            if (g < 0) {
                                for (int m = 0; m < NS; m++) lw[((size_t)j * NS + m) * nCLp + i] = 0.0;
                                continue;
                            }
                            project(c, side ? &p.dx1[(size_t)g * D] : &p.dx0[(size_t)g * D], 0.0, beta, cells);
                            for (int m = 0; m < NS; m++) lw[((size_t)j * NS + m) * nCLp + i] = beta[m];
                        }
                    }
                }
                // The own-cell weight is not stored: the kernel forms it as 1 - (sum of the others), which equals
                // beta[0] when the cell is closed (sum_j Sout_j = 0: a constant field has no gradient).  A mesh
                // whose cells are not closed to round-off cannot use the fused kernel (NaN geometry passes:
                // it is NaN either way).
                auto closed = [&](const double* beta) {
                    double sum = beta[1];
                    for (int m = 2; m < NS; m++) sum += beta[m];
                    const double defect = std::fabs((1.0 - sum) - beta[0]);
                    return !(defect > 1e-12 * (1.0 + std::fabs(beta[0])));
                };
                int nopen = 0;
                for (int lf = 0; lf < nFB; lf++) {
                    const int f = s.flist[lf];
                    const int a = p.fc0[f], b = p.fc1[f];
                    fmeta[lf] = p.meta[f];
                    for (int k = 0; k < D; k++) fSd[(size_t)k * nFBp + lf] = p.Sd[(size_t)f * D + k];
                    if (order != 2) {
                        const int la = local_of(a), lb = b >= 0 ? local_of(b) : 0xFFFF;
                        idx[lf] = (uint32_t)la | ((uint32_t)lb << 16);
                        continue;
                    }
                    double beta[9];
                    int cells[9];
                    // idx rows: [0] = own cells (A | B << 16), [m-1] = stencil entry m >= 2; entry 1 is implicit
                    stencil(a, f, &p.dx0[(size_t)f * D], beta, cells);
                    if (!closed(beta)) nopen++;
                    for (int m = 1; m < NS; m++) w[(size_t)(m - 1) * nFBp + lf] = beta[m];
                    idx[lf] = (uint32_t)cells[0];
                    for (int m = 2; m < NS; m++) idx[(size_t)(m - 1) * nFBp + lf] = (uint32_t)cells[m];
                    if (b >= 0) {
                        stencil(b, f, &p.dx1[(size_t)f * D], beta, cells);
                        if (!closed(beta)) nopen++;
                        for (int m = 1; m < NS; m++) w[(size_t)(NS - 1 + m - 1) * nFBp + lf] = beta[m];
                        idx[lf] |= (uint32_t)cells[0] << 16;
                        for (int m = 2; m < NS; m++) idx[(size_t)(m - 1) * nFBp + lf] |= (uint32_t)cells[m] << 16;
                    } else {
                        for (int m = 1; m < NS; m++) w[(size_t)(NS - 1 + m - 1) * nFBp + lf] = 0.0;
                        idx[lf] |= 0xFFFFu << 16;  // boundary marker
                    }
                }
                if (nopen) {
#pragma omp atomic
                    tp.open_stencils += nopen;
                }
            }
        }
        if (!err.empty()) return err;
    }
    return "";
}

}  // namespace mst
