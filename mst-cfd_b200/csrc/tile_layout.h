// tile_layout.h -- byte layout of one tile of the fused step kernel, shared by
// the host (packing, sizing the launch) and the kernel (carving shared memory).
//
// A tile's read-only tables form ONE contiguous packet in global memory whose
// layout is identical to the first part of the CTA's shared memory, so a single
// TMA bulk copy (cp.async.bulk) stages them.  Work areas follow.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define MST_HD __host__ __device__ __forceinline__
#else
#define MST_HD inline
#endif

namespace mst {

struct TileLayout {
    // packet part (global == shared)
    uint32_t fab;    // u32 [nFXp]      la | lb << 16
    uint32_t feta;   // f64 [nFAp]      (order 2)
    uint32_t fSd;    // f64 [D][nFXp]
    uint32_t slots;  // u16 [nslot][ncgp]
    uint32_t cvol;   // f64 [ncgp]
    uint32_t fdx;    // f64 [2][D][nFBp] (order 2)
    uint32_t fmeta;  // u32 [nFBp]
    uint32_t pk_bytes;
    // work areas (shared only)
    uint32_t mbar, Qs, Rec, Phis, total;
    // padded counts
    uint32_t nFXp, nFAp, nFBp, ncgp, ncg;
};

MST_HD uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }

// order 2: gradient cells = owned + ring 1, faces = all local faces (nFA)
// order 1: gradient cells = owned only (slots for the gather), faces = FB only
MST_HD TileLayout tile_layout(int D, int order, int nslot, int n_own, int n_r1, int n_r2, int nFB, int nFA) {
    const uint32_t U = (uint32_t)D + 2u;
    TileLayout L;
    L.nFAp = (uint32_t)((nFA + 3) & ~3);
    L.nFBp = (uint32_t)((nFB + 3) & ~3);
    L.nFXp = (order == 2) ? L.nFAp : L.nFBp;
    L.ncg = (uint32_t)(order == 2 ? n_own + n_r1 : n_own);
    L.ncgp = (L.ncg + 7u) & ~7u;
    const uint32_t n_loc = (uint32_t)(n_own + n_r1 + n_r2);
    uint32_t o = 0;
    L.fab = o; o += up16(L.nFXp * 4u);
    L.feta = o; if (order == 2) o += up16(L.nFAp * 8u);
    L.fSd = o; o += up16((uint32_t)D * L.nFXp * 8u);
    L.slots = o; o += up16((uint32_t)nslot * L.ncgp * 2u);
    L.cvol = o; o += up16(L.ncgp * 8u);
    L.fdx = o; if (order == 2) o += up16(2u * (uint32_t)D * L.nFBp * 8u);
    L.fmeta = o; o += up16(L.nFBp * 4u);
    L.pk_bytes = o;
    L.mbar = o; o += 16;
    L.Qs = o; o += up16(((n_loc + 1u) & ~1u) * U * 8u);
    L.Rec = o; if (order == 2) o += up16(2u * U * L.nFBp * 8u);
    L.Phis = (order == 2) ? L.Rec : o;  // order 2: Phis overlays Rec[side 0] (same face index, same thread)
    if (order != 2) o += up16(U * L.nFBp * 8u);
    L.total = o;
    return L;
}

}  // namespace mst
