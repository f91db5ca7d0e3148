// plan.h -- host-side layout plan: reference-order mesh tables -> renumbered,
// device-ready tables.  Built once in mstgpu_create().
//
// Input is exactly what the reference's mesh getters expose (include/mstgpu.h,
// mstgpu_mesh).  Output:
//   * cells renumbered along a Morton (Z-order) curve through the cell centres,
//     so that face neighbours are close in memory and any contiguous index
//     range is a compact blob (tiles, partitions);
//   * faces renumbered so that faces of consecutive cells are consecutive
//     (interior faces first, sorted by their lower cell; boundary faces after);
//   * per-cell face lists in ELL form, [slot][cell], keeping each cell's own
//     face ORDER (the reference's file order, R/mesh/MshBlock.cpp:238-239,
//     254-255) so the gather sums in the reference's sequence;
//   * per-face records with everything the flux needs precomputed bit-for-bit
//     as the reference would compute it at run time: Sd = dac*S (outward from
//     c0, MshBlock.cpp:307-318), dx0 = fc - cc[c0], dx1 = fc - cc[c1]
//     (RhoSolver.cpp:250), eta with the off-by-one of RhoSolver.cpp:438 folded
//     in, zone type and left/right flags packed in one word.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mstgpu.h"

namespace mst {

// Lattice of the space-filling curve: origin and extent of the cube its octree subdivides.  Unset = the bounding box of
// the cells being ordered.  How the octree boxes fall on the mesh decides how contiguous a tile's ring cells are in
// memory (on lattice-like meshes runs of consecutive ring ids vary by 40 % with the extent), so the
// extent is chosen by a search (mstgpu.cu, choose_curve_frame) and a partition inherits the frame of the global mesh.
struct CurveFrame {
    bool set = false;
    double lo[3] = {0, 0, 0};
    double ext = 0.0;
};

struct Plan {
    CurveFrame frame;  // INPUT of build_plan (optional)
    int D = 0, U = 0;
    int nc = 0, nf = 0, nint = 0, nslot = 0;
    std::vector<int32_t> cell_new2old, cell_old2new;
    std::vector<int32_t> face_new2old, face_old2new;
    // faces (new order)
    std::vector<int32_t> fc0, fc1;  // new cell ids, fc1 = -1 on boundary faces
    std::vector<double> Sd;         // [nf*D]
    std::vector<double> dx0, dx1;   // [nf*D]
    std::vector<double> eta;        // [nf]   effective eta (1 where Qf = Q[c0])
    std::vector<uint32_t> meta;     // [nf]   type | flags << 8
    // cells (new order)
    std::vector<double> vol;        // [nc]
    std::vector<int32_t> cf;        // [nslot*nc] 2*face + side (side 1: cell is c1), -1 = pad
    // extension (cfg.gradient == MSTGPU_GRAD_LSQ): least-squares gradient as fixed weights,
    //   G_c = sum_j lsq[j][:][c] * (Q_nb(j) - Q_c),   [nslot][D][nc]; zero on boundary / pad slots
    std::vector<double> lsq;
    // extension (cfg.limiter == Venkatakrishnan): eps^2 = K^3 h^3 per cell, [nc]
    std::vector<double> eps2;
};

// returns empty string on success, error text otherwise.
// n_owned >= 0: only cells [0, n_owned) are renumbered (among themselves); the
// rest (ghost cells of a partition) keep their position behind them.
// connectivity checks every host-mesh entry point runs first ("" = fine)
std::string validate_mesh(const mstgpu_mesh& m);
std::string build_plan(const mstgpu_mesh& m, const mstgpu_config& cfg, Plan& p, int n_owned = -1);

// space-filling-curve order of the first n cells (renumber: 1 Morton, 2 Hilbert)
void curve_order(const mstgpu_mesh& m, int renumber, int n, std::vector<int32_t>& new2old, const CurveFrame* frame = nullptr);
// bounding cube of the first n cell centres (the default frame)
CurveFrame bbox_frame(const mstgpu_mesh& m, int n);

}  // namespace mst
