// partition.h -- host-side domain decomposition for multi-GPU runs (one
// partition per GPU).  The reference has no distributed path at all (SURVEY.md
// 2, 8e); this is the build's own layer, kept behind the same C ABI.
//
// A partition's local mesh = its owned cells + `layers` rings of ghost cells
// (2 for the second-order scheme: ghost layer 1 needs its own Green-Gauss
// gradient, which needs layer 2's state; 1 for first order), and every face
// touching an owned or layer-1 cell.  Local cell order: owned cells (ascending
// global id), then ghosts grouped by owner rank (ascending), ascending global
// id inside a group -- so each neighbour's ghosts are one contiguous block
// that ncclRecv can fill in place.  Per-cell face order is preserved, so every
// owned cell sees exactly the arithmetic of the single-domain run: results are
// bit-identical for any partition count.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mstgpu.h"
#include "plan.h"

namespace mst {

struct Neighbor {
    int32_t rank;
    std::vector<int32_t> send_local;  // local ids of owned cells the neighbour holds as ghosts
    int32_t recv_first, recv_count;   // contiguous block of local ghost cells owned by `rank`
};

struct Partition {
    int nparts = 1, rank = 0, layers = 2;
    int D = 0;
    int32_t n_owned = 0, n_ghost1 = 0, n_local = 0, n_faces = 0, n_int = 0;
    std::vector<int32_t> local2global;  // cells
    std::vector<int32_t> face_local2global;
    std::vector<Neighbor> nbrs;
    // local mesh tables (owned by this struct; `mesh` points into them)
    std::vector<int32_t> c0, c1, ftype, cf_ptr, cf_idx;
    std::vector<double> S, fc, eta, cc, vol;
    std::vector<int8_t> dac;
    std::vector<uint8_t> flag;
    mstgpu_mesh mesh{};
    std::vector<int32_t> cell_part;  // [global cells] the assignment this partition was cut from (output path: who owns a node's cells)
    CurveFrame frame;  // curve lattice of the GLOBAL mesh: the local plans order their cells on the same lattice
};

// cell_part: optional [ncells] partition id per global cell; nullptr = split the
// Hilbert-ordered cell list into nparts equal ranges.
std::string build_partition(const mstgpu_mesh& g, const mstgpu_config& cfg, int nparts, int rank,
                            const int32_t* cell_part, Partition& out, const CurveFrame* frame = nullptr);

// the default assignment (Hilbert ranges), exposed for tests / other ranks
std::string default_cell_part(const mstgpu_mesh& g, int nparts, std::vector<int32_t>& part, const CurveFrame* frame = nullptr);

}  // namespace mst
