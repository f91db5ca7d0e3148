// tiles.cpp -- see tiles.h
#include "tiles.h"

#include <algorithm>
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
            int64_t ro = 0, po = 0;
            for (int t = 0; t < nt; t++) {
                TileDesc& d = tp.desc[t];
                const TileLayout L = tile_layout(D, order, nslot, d.n_own, d.n_r1, d.n_r2, d.nFB, d.nFA);
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
                const TileLayout L = tile_layout(D, order, nslot, n_own, n_r1, n_r2, nFB, nFA);
                unsigned char* pk = tp.packets.data() + d.pk_off;
                uint32_t* fab = reinterpret_cast<uint32_t*>(pk + L.fab);
                double* feta = reinterpret_cast<double*>(pk + L.feta);
                double* fSd = reinterpret_cast<double*>(pk + L.fSd);
                uint16_t* slots = reinterpret_cast<uint16_t*>(pk + L.slots);
                double* cvol = reinterpret_cast<double*>(pk + L.cvol);
                double* fdx = reinterpret_cast<double*>(pk + L.fdx);
                uint32_t* fmeta = reinterpret_cast<uint32_t*>(pk + L.fmeta);
                for (size_t i = 0; i < s.ring.size(); i++) tp.ring[d.ring_off + i] = s.ring[i];
                for (uint32_t i = 0; i < (uint32_t)nslot * L.ncgp; i++) slots[i] = 0xFFFF;
                for (uint32_t i = 0; i < L.ncgp; i++) cvol[i] = 1.0;
                for (uint32_t i = 0; i < L.nFXp; i++) fab[i] = 0xFFFFFFFFu;
                for (int lc = 0; lc < (int)L.ncg; lc++) {
                    const int g = lc < n_own ? cb + lc : s.ring[lc - n_own];
                    cvol[lc] = p.vol[g];
                    for (int j = 0; j < nslot; j++) {
                        int f, side;
                        nb_of(g, j, f, side);
                        if (f < 0) continue;
                        const int lf = s.faces.find(f);
                        slots[(size_t)j * L.ncgp + lc] = (uint16_t)((lf << 1) | side);
                    }
                }
                const int nFX = order == 2 ? nFA : nFB;
                for (int lf = 0; lf < nFX; lf++) {
                    const int f = s.flist[lf];
                    const int la = local_of(p.fc0[f]);
                    const int lb = p.fc1[f] >= 0 ? local_of(p.fc1[f]) : 0xFFFF;
                    fab[lf] = (uint32_t)(la < 0 ? 0xFFFF : la) | ((uint32_t)(lb < 0 ? 0xFFFF : lb) << 16);
                    if (order == 2) feta[lf] = p.eta[f];
                    for (int k = 0; k < D; k++) fSd[(size_t)k * L.nFXp + lf] = p.Sd[(size_t)f * D + k];
                    if (lf < nFB) {
                        fmeta[lf] = p.meta[f];
                        if (order == 2)
                            for (int k = 0; k < D; k++) {
                                fdx[(size_t)k * L.nFBp + lf] = p.dx0[(size_t)f * D + k];
                                fdx[(size_t)(D + k) * L.nFBp + lf] = p.dx1[(size_t)f * D + k];
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
