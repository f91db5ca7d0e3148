// partition.cpp -- see partition.h
#include "partition.h"

#include <algorithm>
#include <numeric>

#include "plan.h"

namespace mst {

std::string default_cell_part(const mstgpu_mesh& g, int nparts, std::vector<int32_t>& part, const CurveFrame* frame) {
    if (nparts < 1) return "nparts < 1";
    std::vector<int32_t> ord;
    curve_order(g, 2, g.ncells, ord, frame);
    part.resize(g.ncells);
    // equal ranges of the Hilbert curve: compact, balanced to +-1 cell
    for (int64_t i = 0; i < g.ncells; i++) part[ord[i]] = (int32_t)(i * nparts / g.ncells);
    return "";
}

std::string build_partition(const mstgpu_mesh& g, const mstgpu_config& cfg, int nparts, int rank,
                            const int32_t* cell_part, Partition& P, const CurveFrame* frame) {
    {
        std::string verr = validate_mesh(g);
        if (!verr.empty()) return verr;
    }
    const int D = g.dim, nc = g.ncells, nf = g.nfaces;
    if (nparts < 1 || rank < 0 || rank >= nparts) return "bad nparts / rank";
    std::vector<int32_t> own_part;
    if (!cell_part) {
        std::string e = default_cell_part(g, nparts, own_part, frame);
        if (!e.empty()) return e;
        cell_part = own_part.data();
    }
    P.nparts = nparts; P.rank = rank; P.D = D;
    P.cell_part.assign(cell_part, cell_part + nc);
    P.frame = (frame && frame->set) ? *frame : bbox_frame(g, nc);
    // the viscous term needs the primitive gradient of layer-1 cells as well
    P.layers = (cfg.order == 2 || cfg.viscous != 0) ? 2 : 1;
    const int L = P.layers;

    auto other = [&](int f, int c) -> int { return g.c0[f] == c ? g.c1[f] : g.c0[f]; };

    // ---- ghost layers by face adjacency ---------------------------------------------
    std::vector<int8_t> lay(nc, -1);
    std::vector<int32_t> owned, front, ghosts;
    for (int c = 0; c < nc; c++)
        if (cell_part[c] == rank) { lay[c] = 0; owned.push_back(c); }
    if (owned.empty()) return "partition owns no cells";
    front = owned;
    std::vector<int32_t> layer_count(L + 1, 0);
    for (int l = 1; l <= L; l++) {
        std::vector<int32_t> next;
        for (int c : front)
            for (int j = g.cf_ptr[c]; j < g.cf_ptr[c + 1]; j++) {
                const int nb = other(g.cf_idx[j], c);
                if (nb >= 0 && lay[nb] < 0) { lay[nb] = (int8_t)l; next.push_back(nb); }
            }
        layer_count[l] = (int32_t)next.size();
        ghosts.insert(ghosts.end(), next.begin(), next.end());
        front.swap(next);
    }
    // ghosts grouped by owner rank, ascending global id inside a group
    std::sort(ghosts.begin(), ghosts.end(), [&](int a, int b) {
        return cell_part[a] != cell_part[b] ? cell_part[a] < cell_part[b] : a < b;
    });
    P.n_owned = (int32_t)owned.size();
    P.n_ghost1 = layer_count[1];
    P.n_local = P.n_owned + (int32_t)ghosts.size();
    P.local2global = owned;
    P.local2global.insert(P.local2global.end(), ghosts.begin(), ghosts.end());
    std::vector<int32_t> g2l(nc, -1);
    for (int i = 0; i < P.n_local; i++) g2l[P.local2global[i]] = i;

    // ---- faces: every face touching an owned or inner-layer cell ---------------------
    // (a cell of the outermost layer only lends its state, its other faces are dropped)
    std::vector<int32_t> faces;
    for (int f = 0; f < nf; f++) {
        const int a = g.c0[f], b = g.c1[f];
        const bool ina = lay[a] >= 0 && lay[a] < L;
        const bool inb = b >= 0 && lay[b] >= 0 && lay[b] < L;
        if (ina || inb) faces.push_back(f);
    }
    // interior faces first, global order inside each group
    std::stable_sort(faces.begin(), faces.end(), [&](int a, int b) {
        const bool ia = g.c1[a] >= 0 && g.ftype[a] == MSTGPU_BC_INTERIOR;
        const bool ib = g.c1[b] >= 0 && g.ftype[b] == MSTGPU_BC_INTERIOR;
        return ia != ib ? ia : false;
    });
    P.n_faces = (int32_t)faces.size();
    P.face_local2global = faces;
    std::vector<int32_t> fg2l(nf, -1);
    for (int i = 0; i < P.n_faces; i++) fg2l[faces[i]] = i;

    const int qf_from = cfg.qf_copy_from < 0 ? g.nint - 1 : cfg.qf_copy_from;
    P.c0.resize(P.n_faces); P.c1.resize(P.n_faces); P.ftype.resize(P.n_faces);
    P.S.resize((size_t)P.n_faces * D); P.fc.resize((size_t)P.n_faces * D); P.eta.resize(P.n_faces);
    P.dac.resize(P.n_faces); P.flag.resize((size_t)P.n_faces * D);
    P.n_int = 0;
    for (int i = 0; i < P.n_faces; i++) {
        const int f = faces[i];
        const bool interior = g.c1[f] >= 0 && g.ftype[f] == MSTGPU_BC_INTERIOR;
        P.c0[i] = g2l[g.c0[f]];
        P.c1[i] = interior ? g2l[g.c1[f]] : -1;
        if (P.c0[i] < 0 || (interior && P.c1[i] < 0)) return "internal: face with a cell outside the local mesh";
        P.ftype[i] = g.ftype[f];
        P.dac[i] = g.dac[f];
        // the off-by-one of RhoSolver.cpp:438 is defined on GLOBAL face ids: bake it in
        P.eta[i] = (f >= qf_from || !interior) ? 1.0 : g.eta[f];
        for (int d = 0; d < D; d++) {
            P.S[(size_t)i * D + d] = g.S[(size_t)f * D + d];
            P.fc[(size_t)i * D + d] = g.fc[(size_t)f * D + d];
            P.flag[(size_t)i * D + d] = g.flag[(size_t)f * D + d];
        }
        if (interior) P.n_int++;
    }
    // ---- cells -------------------------------------------------------------------------
    P.cc.resize((size_t)P.n_local * D); P.vol.resize(P.n_local); P.cf_ptr.assign(P.n_local + 1, 0);
    for (int i = 0; i < P.n_local; i++) {
        const int c = P.local2global[i];
        P.vol[i] = g.vol[c];
        for (int d = 0; d < D; d++) P.cc[(size_t)i * D + d] = g.cc[(size_t)c * D + d];
        for (int j = g.cf_ptr[c]; j < g.cf_ptr[c + 1]; j++) {
            const int lf = fg2l[g.cf_idx[j]];
            if (lf >= 0) P.cf_idx.push_back(lf);  // the cell's own (file) order is kept
        }
        P.cf_ptr[i + 1] = (int32_t)P.cf_idx.size();
    }
    // ---- neighbours ----------------------------------------------------------------------
    P.nbrs.clear();
    for (int i = P.n_owned; i < P.n_local;) {
        const int r = cell_part[P.local2global[i]];
        int j = i;
        while (j < P.n_local && cell_part[P.local2global[j]] == r) j++;
        Neighbor nb;
        nb.rank = r; nb.recv_first = i; nb.recv_count = j - i;
        P.nbrs.push_back(nb);
        i = j;
    }
    // send lists: owned cell c is a ghost of rank s iff a cell of s lies within L faces of c
    {
        std::vector<std::vector<int32_t>> send(nparts);
        std::vector<int32_t> seen_r;
        std::vector<int32_t> a, b;
        for (int i = 0; i < P.n_owned; i++) {
            const int c = owned[i];
            // quick reject: a cell whose 2-neighbourhood is all ours (lay == 0 everywhere)
            seen_r.clear();
            a.assign(1, c);
            for (int l = 1; l <= L; l++) {
                b.clear();
                for (int x : a)
                    for (int j = g.cf_ptr[x]; j < g.cf_ptr[x + 1]; j++) {
                        const int nbc = other(g.cf_idx[j], x);
                        if (nbc < 0) continue;
                        const int r = cell_part[nbc];
                        if (r != rank && std::find(seen_r.begin(), seen_r.end(), r) == seen_r.end()) seen_r.push_back(r);
                        b.push_back(nbc);
                    }
                a.swap(b);
            }
            for (int r : seen_r) send[r].push_back(i);
        }
        for (int s = 0; s < nparts; s++) {
            if (send[s].empty()) continue;
            auto it = std::find_if(P.nbrs.begin(), P.nbrs.end(), [&](const Neighbor& n) { return n.rank == s; });
            if (it == P.nbrs.end()) {  // we send to s but receive nothing from it (cannot happen on a symmetric graph)
                Neighbor nb;
                nb.rank = s; nb.recv_first = P.n_local; nb.recv_count = 0;
                P.nbrs.push_back(nb);
                it = P.nbrs.end() - 1;
            }
            it->send_local = send[s];
        }
        std::sort(P.nbrs.begin(), P.nbrs.end(), [](const Neighbor& x, const Neighbor& y) { return x.rank < y.rank; });
    }
    mstgpu_mesh& m = P.mesh;
    m.dim = D; m.ncells = P.n_local; m.nfaces = P.n_faces; m.nint = P.n_int;
    m.c0 = P.c0.data(); m.c1 = P.c1.data(); m.S = P.S.data(); m.dac = P.dac.data(); m.fc = P.fc.data();
    m.eta = P.eta.data(); m.flag = P.flag.data(); m.ftype = P.ftype.data(); m.cc = P.cc.data();
    m.vol = P.vol.data(); m.cf_ptr = P.cf_ptr.data(); m.cf_idx = P.cf_idx.data();
    return "";
}

}  // namespace mst
