// tiles.h -- tile packets for the fused step kernel (k_step_tiles).
//
// The device-order cell range [0, n_update) is cut into tiles of T consecutive
// cells (Morton order makes a tile a compact blob).  One CTA advances one tile
// by a full time step without touching HBM for intermediates:
//
//   owned cells   the T cells the tile updates
//   ring 1        face neighbours of owned cells outside the tile: their state
//                 AND their gradient are needed (2nd-order reconstruction on
//                 the tile's outer faces) -> gradient recomputed redundantly
//   ring 2        face neighbours of ring-1 cells: state only (order 2)
//   FB faces      faces with an owned cell: flux evaluated here
//   FG faces      remaining faces of ring-1 cells: geometry for their gradient
//
// Everything a tile needs is stored in ITS OWN contiguous packet (16-byte
// aligned sub-arrays, 16-bit local indices, layout in tile_layout.h), so the
// kernel stages it with ONE TMA bulk copy instead of scattered gathers.
// Reference semantics are unchanged: per-cell face order, eta, Sd, dx0/dx1,
// flags are the values of plan.h, merely regrouped.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "plan.h"

namespace mst {

struct TileDesc {
    int32_t cb;     // first owned cell (device order)
    int32_t n_own;  // owned cells
    int32_t n_r1;   // ring-1 cells
    int32_t n_r2;   // ring-2 cells
    int32_t nFB;    // flux faces
    int32_t nFA;    // all local faces (FB first)
    int64_t ring_off;  // into ring[] (n_r1 + n_r2 entries, padded to 4)
    int64_t pk_off;    // byte offset of the packet in packets[] (16-byte aligned)
};

struct TilePack {
    int T = 0, ntiles = 0, order = 2, D = 0, nslot = 0, ext = 0;
    std::vector<TileDesc> desc;
    std::vector<int32_t> ring;           // device-order cell ids of ring cells
    std::vector<unsigned char> packets;  // concatenated packets, layout = tile_layout()
    size_t max_smem = 0;                 // dynamic shared memory the largest tile needs
    int64_t sum_r1 = 0, sum_r2 = 0, sum_FB = 0, sum_FA = 0;
    int64_t open_stencils = 0;           // face sides whose cell is not closed (own-cell weight != 1 - sum of the others)
};

// n_update: cells [0, n_update) are advanced (the rest are ghosts, read only)
// ext (tile_ext()): the packets also carry the limiter / viscous tables of tile_layout.h
// fit_faces > 0: variable tile sizes, each grown until its flux faces would exceed fit_faces (or its cells T)
std::string build_tiles(const Plan& p, int n_update, int T, int order, TilePack& tp, int ext = 0, int fit_faces = 0);

}  // namespace mst
