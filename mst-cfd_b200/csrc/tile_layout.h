// tile_layout.h -- byte layout of one tile of the fused step kernel, shared by
// the host (packing, sizing the launch) and the kernel.
//
// A tile's read-only tables form one contiguous packet in global memory.  Every
// array in it is read exactly once per step by the thread that owns the face /
// cell (SoA, coalesced), so the packet is streamed straight from HBM into
// registers; shared memory only holds what is gathered at random: the cell
// states of the tile and its rings (Qs) and the face fluxes (Phis).
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
    // packet (global memory), byte offsets from the packet start
    uint32_t w;      // f64 [2*(NS-1)][nFBp]  reconstruction weights of stencil entries 1..NS-1: side A (c0) then side B
                     //                   (c1) (order 2).  The own-cell weight is NOT stored: a closed cell reproduces
                     //                   constants, so it is 1 - sum of the others (checked at packing time)
    uint32_t idx;    // u32 [NS-1][nFBp]  local cell ids, A | B << 16: row 0 = the face's own cells (la | lb),
                     //                   row m-1 = stencil entry m >= 2 (entry 1 is the cell across the face)
    uint32_t fSd;    // f64 [D][nFBp]     area vector, outward from c0
    uint32_t fmeta;  // u32 [nFBp]        zone type | left/right flags << 8
    // extensions (ext bit 0: limiter, bit 1: viscous term): per cell of the tile and of ring 1
    // ("stencil cells", local ids 0..nCL-1)
    uint32_t lw;     // f64 [nslot][NS][nCLp]  (limiter) slope at face centre j:  D_j = sum_m lw[j][m] Q_stencil(m)
    uint32_t lid;    // u16 [nslot][nCLp]      local ids of the face neighbours (own id where there is none)
    uint32_t le2;    // f64 [nCLp]             (limiter) Venkatakrishnan eps^2
    uint32_t vw;     // f64 [nslot][2+D][nCLp] (viscous) per face slot: weight of the cell, weight of the neighbour
                     //                        in the face state, outward area vector / V
    uint32_t feta;   // f64 [nFBp]             (viscous) eta of the flux face (1 where the face state is Q[c0])
    uint32_t slots;  // u16 [nslot][ncp]  per owned cell: local face << 1 | side, 0xFFFF = pad
    uint32_t cvol;   // f64 [ncp]         1 / cell volume
    uint32_t pk_bytes;
    // shared memory
    uint32_t mbar, Qs, Phis, cells_s, total;  // cells_s: copy of the packet's slots + cvol block
    uint32_t stage;                           // (ext & 4) per-thread landing slots of the NEXT face iteration's weights and
                                              // stencil ids: f64 [2*(NS-1)][NT] then u32 [NS-1][NT], filled by cp.async
    uint32_t philim;                          // f64 [U][nCLp] limiter value per stencil cell (ext & 1)
    uint32_t Gps;                             // f64 [(D+1)*D][nCLp] primitive gradients (u_i, T) per stencil cell (ext & 2)
    uint32_t nFBp, ncp, nCLp;
};

MST_HD uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }

// NS = stencil size of the second-order reconstruction = 1 + max faces per cell
// ext: bit 0 = limiter tables, bit 1 = viscous tables (both need order == 2: rings 1 and 2);
//      bit 2 = staged packet stream (shared memory only, the packet is the same), CTA size in bits 8..
MST_HD int tile_ext(int order, int limiter, int viscous) { return order == 2 ? ((limiter != 0 ? 1 : 0) | (viscous != 0 ? 2 : 0)) : 0; }
MST_HD int tile_ext_staged(int ext, int nthreads) { return ext | 4 | (nthreads << 8); }

MST_HD TileLayout tile_layout(int D, int order, int nslot, int n_own, int n_r1, int n_r2, int nFB, int ext = 0) {
    const uint32_t U = (uint32_t)D + 2u;
    const uint32_t NS = (order == 2) ? (uint32_t)nslot + 1u : 1u;
    TileLayout L;
    L.nFBp = (uint32_t)((nFB + 3) & ~3);
    L.ncp = ((uint32_t)n_own + 7u) & ~7u;
    const uint32_t n_loc = (uint32_t)(n_own + n_r1 + n_r2);
    uint32_t o = 0;
    L.w = o; if (order == 2) o += up16(2u * (NS - 1u) * L.nFBp * 8u);
    L.idx = o; o += up16((NS > 1u ? NS - 1u : 1u) * L.nFBp * 4u);
    L.fSd = o; o += up16((uint32_t)D * L.nFBp * 8u);
    L.fmeta = o; o += up16(L.nFBp * 4u);
    if (order != 2) ext = 0;
    const bool limited = (ext & 1) != 0, visc = (ext & 2) != 0;
    L.nCLp = ext ? (((uint32_t)(n_own + n_r1) + 3u) & ~3u) : 0u;
    L.lw = o; if (limited) o += up16((uint32_t)nslot * NS * L.nCLp * 8u);
    L.lid = o; o += up16((uint32_t)nslot * L.nCLp * 2u);
    L.le2 = o; if (limited) o += up16(L.nCLp * 8u);
    L.vw = o; if (visc) o += up16((uint32_t)nslot * (2u + (uint32_t)D) * L.nCLp * 8u);
    L.feta = o; if (visc) o += up16(L.nFBp * 8u);
    L.slots = o; o += up16((uint32_t)nslot * L.ncp * 2u);
    L.cvol = o; o += up16(L.ncp * 8u);
    L.pk_bytes = o;
    uint32_t s = 0;
    L.mbar = s; s += 16;
    L.Qs = s; s += up16(((n_loc + 1u) & ~1u) * U * 8u);
    L.Phis = s; s += up16(U * L.nFBp * 8u);
    L.cells_s = s; s += L.pk_bytes - L.slots;
    L.philim = s; if (limited) s += up16(U * L.nCLp * 8u);
    L.Gps = s; if (visc) s += up16(((uint32_t)D + 1u) * (uint32_t)D * L.nCLp * 8u);
    L.stage = s; if (ext & 4) s += up16(((uint32_t)ext >> 8) * (2u * (NS - 1u) * 8u + (NS - 1u) * 4u));
    L.total = s;
    return L;
}

}  // namespace mst
