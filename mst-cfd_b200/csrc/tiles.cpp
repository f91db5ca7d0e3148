// tiles.cpp -- see tiles.h
#include "tiles.h"

#include <algorithm>
#include <cstring>

#include "tile_layout.h"

namespace mst {

size_t tile_smem_bytes(int D, int order, int n_own, int n_r1, int n_r2, int nFB, int nFA) {
    return tile_layout(D, order, n_own, n_r1, n_r2, nFB, nFA).total;
}

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
    std::vector<int32_t> ring, flist;   // ring cell ids; local face -> device face id
};

inline int up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

std::string build_tiles(const Plan& p, int n_update, int T, int order, TilePack& tp) {
    const int D = p.D, nc = p.nc, nslot = p.nslot;
    if (T < 32 || T > 2048 || (T & 1)) return "tile size must be even, 32..2048";
    if (n_update <= 0 || n_update > nc) return "n_update out of range";
    tp.T = T; tp.order = order; tp.D = D; tp.nslot = nslot;
    tp.ntiles = (n_update + T - 1) / T;
    tp.desc.assign(tp.ntiles, TileDesc{});
    const int nt = tp.ntiles;

    auto nb_of = [&](int c, int j, int& f, int& side) -> int {
        const int v = p.cf[(size_t)j * nc + c];
        if (v < 0) { f = -1; side = 0; return -2; }
        f = v >> 1; side = v & 1;
        return side ? p.fc0[f] : p.fc1[f];
    };

    // pass 1: sizes; pass 2: fill.  The traversal is identical in both passes.
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int64_t ro = 0, co = 0, fa = 0, fb = 0;
            for (int t = 0; t < nt; t++) {
                TileDesc& d = tp.desc[t];
                d.ring_off = ro; d.cell_off = co; d.fa_off = fa; d.fb_off = fb;
                ro += up(d.n_r1 + d.n_r2, 4);
                co += up(d.n_own + d.n_r1, 8);
                fa += up(d.nFA, 4);
                fb += up(d.nFB, 4);
                tp.max_smem = std::max(tp.max_smem, tile_smem_bytes(D, order, d.n_own, d.n_r1, d.n_r2, d.nFB, d.nFA));
                tp.sum_r1 += d.n_r1; tp.sum_r2 += d.n_r2; tp.sum_FB += d.nFB; tp.sum_FA += d.nFA;
            }
            tp.ring.assign(ro, 0);
            tp.slots.assign((size_t)nslot * co, 0xFFFF);
            tp.cvol.assign(co, 1.0);
            tp.fab.assign(fa, 0xFFFFFFFFu);
            tp.feta.assign(fa, 1.0);
            tp.fSd.assign((size_t)D * fa, 0.0);
            tp.fdx.assign((size_t)2 * D * fb, 0.0);
            tp.fmeta.assign(fb, 0);
        }
        std::string err;
#pragma omp parallel
        {
            TileScratch s;
#pragma omp for schedule(dynamic, 64)
            for (int t = 0; t < nt; t++) {
                TileDesc& d = tp.desc[t];
                const int cb = t * T, ce = std::min(cb + T, n_update), n_own = ce - cb;
                s.cells.reset(); s.faces.reset(); s.ring.clear(); s.flist.clear();
                auto local_of = [&](int g) -> int { return (g >= cb && g < ce) ? g - cb : s.cells.find(g); };
                int n_r1 = 0, n_r2 = 0;
                for (int c = cb; c < ce; c++)
                    for (int j = 0; j < nslot; j++) {
                        int f, side;
                        const int nb = nb_of(c, j, f, side);
                        if (nb >= 0 && !(nb >= cb && nb < ce) && s.cells.find(nb) < 0) {
                            s.cells.put(nb, n_own + n_r1++);
                            s.ring.push_back(nb);
                        }
                    }
                if (order == 2)
                    for (int i = 0; i < n_r1; i++) {
                        const int r = s.ring[i];
                        for (int j = 0; j < nslot; j++) {
                            int f, side;
                            const int nb = nb_of(r, j, f, side);
                            if (nb >= 0 && !(nb >= cb && nb < ce) && s.cells.find(nb) < 0) {
                                s.cells.put(nb, n_own + n_r1 + n_r2++);
                                s.ring.push_back(nb);
                            }
                        }
                    }
                // FB: faces of owned cells, each once
                for (int c = cb; c < ce; c++)
                    for (int j = 0; j < nslot; j++) {
                        int f, side;
                        const int nb = nb_of(c, j, f, side);
                        if (f < 0) continue;
                        const bool nb_owned = nb >= cb && nb < ce;
                        if (nb_owned && side == 1) continue;  // the c0-side cell lists it
                        if (s.faces.find(f) >= 0) continue;
                        s.faces.put(f, (int)s.flist.size());
                        s.flist.push_back(f);
                    }
                const int nFB = (int)s.flist.size();
                if (order == 2)
                    for (int i = 0; i < n_r1; i++) {
                        const int r = s.ring[i];
                        for (int j = 0; j < nslot; j++) {
                            int f, side;
                            nb_of(r, j, f, side);
                            if (f < 0 || s.faces.find(f) >= 0) continue;
                            s.faces.put(f, (int)s.flist.size());
                            s.flist.push_back(f);
                        }
                    }
                const int nFA = (int)s.flist.size();
                if (pass == 0) {
                    d.cb = cb; d.n_own = n_own; d.n_r1 = n_r1; d.n_r2 = n_r2; d.nFB = nFB; d.nFA = nFA;
                    if (n_own + n_r1 + n_r2 >= 0xFFFF || nFA >= 0x7FFF || n_own + n_r1 + n_r2 > 6000 || nFA > 6000) {
#pragma omp critical
                        err = "tile too large for 16-bit local indices";
                    }
                    continue;
                }
                // ---- fill ---------------------------------------------------------
                const int ncg = n_own + n_r1, ncgp = up(ncg, 8), nFAp = up(nFA, 4), nFBp = up(nFB, 4);
                for (size_t i = 0; i < s.ring.size(); i++) tp.ring[d.ring_off + i] = s.ring[i];
                for (int lc = 0; lc < ncg; lc++) {
                    const int g = lc < n_own ? cb + lc : s.ring[lc - n_own];
                    tp.cvol[d.cell_off + lc] = p.vol[g];
                    for (int j = 0; j < nslot; j++) {
                        int f, side;
                        nb_of(g, j, f, side);
                        if (f < 0) continue;
                        const int lf = s.faces.find(f);
                        tp.slots[(size_t)nslot * d.cell_off + (size_t)j * ncgp + lc] = (uint16_t)((lf << 1) | side);
                    }
                }
                for (int lf = 0; lf < nFA; lf++) {
                    const int f = s.flist[lf];
                    const int la = local_of(p.fc0[f]);
                    const int lb = p.fc1[f] >= 0 ? local_of(p.fc1[f]) : 0xFFFF;
                    // a ring-2/outside cell can be absent only on faces no gradient cell uses from that side
                    tp.fab[d.fa_off + lf] = (uint32_t)(la < 0 ? 0xFFFF : la) | ((uint32_t)(lb < 0 ? 0xFFFF : lb) << 16);
                    tp.feta[d.fa_off + lf] = p.eta[f];
                    for (int k = 0; k < D; k++) tp.fSd[(size_t)D * d.fa_off + (size_t)k * nFAp + lf] = p.Sd[(size_t)f * D + k];
                    if (lf < nFB) {
                        tp.fmeta[d.fb_off + lf] = p.meta[f];
                        for (int k = 0; k < D; k++) {
                            tp.fdx[(size_t)2 * D * d.fb_off + (size_t)k * nFBp + lf] = p.dx0[(size_t)f * D + k];
                            tp.fdx[(size_t)2 * D * d.fb_off + (size_t)(D + k) * nFBp + lf] = p.dx1[(size_t)f * D + k];
                        }
                    }
                }
            }
        }
        if (!err.empty()) return err;
    }
    return "";
}

}  // namespace mst
