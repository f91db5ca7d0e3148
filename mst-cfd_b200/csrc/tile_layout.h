// tile_layout.h -- shared-memory carve-up of k_step_tiles, shared by the host
// (sizing the launch) and the kernel (computing the pointers).
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define MST_HD __host__ __device__ __forceinline__
#else
#define MST_HD inline
#endif

namespace mst {

struct TileSmem {
    uint32_t mbar, Qs, Gs, Phis, Qout, fab, feta, fSd, total;
};

MST_HD uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }

MST_HD TileSmem tile_layout(int D, int order, int n_own, int n_r1, int n_r2, int nFB, int nFA) {
    const int U = D + 2;
    const uint32_t n_loc = (uint32_t)(n_own + n_r1 + n_r2);
    const uint32_t ncg = (uint32_t)(n_own + n_r1);
    const uint32_t nFAp = (uint32_t)((nFA + 3) & ~3);
    TileSmem s;
    uint32_t o = 0;
    s.mbar = o; o += 16;
    s.Qs = o; o += up16(((n_loc + 1u) & ~1u) * U * 8u);
    s.Gs = o; if (order == 2) o += up16(ncg * U * D * 8u);
    s.Phis = o; o += up16((uint32_t)nFB * U * 8u);
    s.Qout = o; o += up16((((uint32_t)n_own + 1u) & ~1u) * U * 8u);
    s.fab = o; if (order == 2) o += up16(nFAp * 4u);
    s.feta = o; if (order == 2) o += up16(nFAp * 8u);
    s.fSd = o; if (order == 2) o += up16((uint32_t)D * nFAp * 8u);
    s.total = o;
    return s;
}

}  // namespace mst
