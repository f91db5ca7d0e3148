// step_tiles.cuh -- the fused step kernel: one CTA advances one tile of cells by
// a full explicit step (gradient -> reconstruction -> Roe/AUSM+ flux -> gather ->
// Euler update -> residual) with every intermediate in shared memory.
//
// Replaces, in one launch, the reference loops
//   R/rhoSolver/RhoSolver.cpp:430-452  (face interpolation + Green-Gauss)
//   R/rhoSolver/RhoSolver.cpp:90-369   (reconstruction + per-direction Riemann flux, BCs)
//   R/rhoSolver/RhoSolver.cpp:45-68    (cell gather + explicit Euler)
//   R/time/Time.cpp:69-76              (L-inf residual)
// HBM traffic per step: the tile packet (geometry, streamed once), Q of the
// tile and its rings in, Q of the tile out.  Neither the cell gradients
// (120 B/cell) nor the face fluxes (2 x 40 B/cell) ever reach HBM.
//
// sm_100a specifics: the packet arrays and the owned block of Q are staged by
// TMA bulk copies (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier; the
// new state leaves through a bulk store.  Ring cells are 40-byte row gathers.
#pragma once
#include <cstdint>

#include "physics.cuh"
#include "tile_layout.h"
#include "tiles.h"

namespace mst {

struct TileArrays {
    const TileDesc* desc;
    const int32_t* ring;
    const uint16_t* slots;
    const double* cvol;
    const uint32_t* fab;
    const double* feta;
    const double* fSd;
    const double* fdx;
    const uint32_t* fmeta;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ double warp_max_d(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

template <int D, int ORDER, int NT>
__global__ void __launch_bounds__(NT) k_step_tiles(TileArrays ta, DevCfg cfg, int nslot, double dt,
                                                   const double* __restrict__ Qold,
                                                   double* __restrict__ Qnew,
                                                   unsigned long long* __restrict__ resid,
                                                   int* __restrict__ nanflag) {
    constexpr int U = D + 2;
    extern __shared__ __align__(128) unsigned char smem[];
    const TileDesc d = ta.desc[blockIdx.x];
    const int tid = threadIdx.x;
    const int n_own = d.n_own, n_ring = d.n_r1 + d.n_r2, ncg = d.n_own + d.n_r1;
    const int nFB = d.nFB, nFAp = (d.nFA + 3) & ~3, nFBp = (d.nFB + 3) & ~3, ncgp = (ncg + 7) & ~7;
    const TileSmem L = tile_layout(D, ORDER, d.n_own, d.n_r1, d.n_r2, d.nFB, d.nFA);
    double* Qs = reinterpret_cast<double*>(smem + L.Qs);
    double* Gs = reinterpret_cast<double*>(smem + L.Gs);
    double* Phis = reinterpret_cast<double*>(smem + L.Phis);
    double* Qout = reinterpret_cast<double*>(smem + L.Qout);
    const uint32_t* fab_s = reinterpret_cast<const uint32_t*>(smem + L.fab);
    const double* feta_s = reinterpret_cast<const double*>(smem + L.feta);
    const double* fSd_s = reinterpret_cast<const double*>(smem + L.fSd);
    const uint32_t bar = smem_u32(smem + L.mbar);

    // ---- phase 0: stage the tile ------------------------------------------------
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t qbytes = (uint32_t)((n_own + 1) & ~1) * U * 8u;
        uint32_t tx = qbytes;
        if (ORDER == 2) tx += (uint32_t)nFAp * 4u + (uint32_t)nFAp * 8u + (uint32_t)D * nFAp * 8u;
        mbar_expect_tx(bar, tx);
        bulk_g2s(smem_u32(Qs), Qold + (size_t)d.cb * U, qbytes, bar);
        if (ORDER == 2) {
            bulk_g2s(smem_u32(fab_s), ta.fab + d.fa_off, (uint32_t)nFAp * 4u, bar);
            bulk_g2s(smem_u32(feta_s), ta.feta + d.fa_off, (uint32_t)nFAp * 8u, bar);
            bulk_g2s(smem_u32(fSd_s), ta.fSd + (size_t)D * d.fa_off, (uint32_t)D * nFAp * 8u, bar);
        }
    }
    // ring cells: row gathers (rows are contiguous, 8*U bytes)
    for (int i = tid; i < n_ring * U; i += NT) {
        const int r = i / U, k = i - r * U;
        const int g = ta.ring[d.ring_off + r];
        Qs[(size_t)(n_own + r) * U + k] = Qold[(size_t)g * U + k];
    }
    mbar_wait(bar, 0);
    __syncthreads();

    // ---- phase 1: Green-Gauss gradients of owned + ring-1 cells --------------------
    if (ORDER == 2) {
        const uint16_t* slots = ta.slots + (size_t)nslot * d.cell_off;
        for (int lc = tid; lc < ncg; lc += NT) {
            double qc[U];
#pragma unroll
            for (int k = 0; k < U; k++) qc[k] = Qs[lc * U + k];
            double t[U][D];
#pragma unroll
            for (int k = 0; k < U; k++)
#pragma unroll
                for (int dd = 0; dd < D; dd++) t[k][dd] = 0.0;
            for (int j = 0; j < nslot; j++) {
                const uint32_t v = slots[(size_t)j * ncgp + lc];
                if (v == 0xFFFFu) continue;
                const int lf = v >> 1;
                const int side = v & 1;
                const uint32_t ab = fab_s[lf];
                const uint32_t nb = side ? (ab & 0xFFFFu) : (ab >> 16);
                const double e = feta_s[lf];
                double qf[U];
                if (nb != 0xFFFFu) {
                    const double e0 = side ? (1.0 - e) : e;
                    const double e1 = side ? e : (1.0 - e);
#pragma unroll
                    for (int k = 0; k < U; k++) qf[k] = e0 * qc[k] + e1 * Qs[nb * U + k];
                } else {
#pragma unroll
                    for (int k = 0; k < U; k++) qf[k] = qc[k];
                }
                const double sg = side ? -1.0 : 1.0;
#pragma unroll
                for (int dd = 0; dd < D; dd++) {
                    const double s = sg * fSd_s[dd * nFAp + lf];
#pragma unroll
                    for (int k = 0; k < U; k++) t[k][dd] += qf[k] * s;
                }
            }
            const double v = ta.cvol[d.cell_off + lc];
#pragma unroll
            for (int k = 0; k < U; k++)
#pragma unroll
                for (int dd = 0; dd < D; dd++) Gs[(lc * U + k) * D + dd] = t[k][dd] / v;
        }
        __syncthreads();
    }

    // ---- phase 2: reconstruction + flux on every face with an owned cell ----------
    {
        const double* dxg = ta.fdx + (size_t)2 * D * d.fb_off;
        for (int f = tid; f < nFB; f += NT) {
            const uint32_t ab = (ORDER == 2) ? fab_s[f] : ta.fab[d.fa_off + f];
            const int la = ab & 0xFFFFu, lb = ab >> 16;
            const uint32_t mt = ta.fmeta[d.fb_off + f];
            const int type = mt & 0xff;
            const uint32_t flags = mt >> 8;
            double S[D];
#pragma unroll
            for (int dd = 0; dd < D; dd++)
                S[dd] = (ORDER == 2) ? fSd_s[dd * nFAp + f] : ta.fSd[(size_t)D * d.fa_off + (size_t)dd * nFAp + f];
            double qa[U], ra[U];
#pragma unroll
            for (int k = 0; k < U; k++) qa[k] = Qs[la * U + k];
            if (ORDER == 2) {
                double dx[D];
#pragma unroll
                for (int dd = 0; dd < D; dd++) dx[dd] = dxg[(size_t)dd * nFBp + f];
#pragma unroll
                for (int k = 0; k < U; k++) {
                    double s = 0.0;
#pragma unroll
                    for (int dd = 0; dd < D; dd++) s += Gs[(la * U + k) * D + dd] * dx[dd];
                    ra[k] = qa[k] + s;
                }
            } else {
#pragma unroll
                for (int k = 0; k < U; k++) ra[k] = qa[k];
            }
            double A[U], B[U], phi[U];
            bool live = true;
            if (lb != 0xFFFF) {
#pragma unroll
                for (int k = 0; k < U; k++) A[k] = ra[k];
                if (ORDER == 2) {
                    double dx[D];
#pragma unroll
                    for (int dd = 0; dd < D; dd++) dx[dd] = dxg[(size_t)(D + dd) * nFBp + f];
#pragma unroll
                    for (int k = 0; k < U; k++) {
                        double s = 0.0;
#pragma unroll
                        for (int dd = 0; dd < D; dd++) s += Gs[(lb * U + k) * D + dd] * dx[dd];
                        B[k] = Qs[lb * U + k] + s;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < U; k++) B[k] = Qs[lb * U + k];
                }
            } else {
                live = boundary_states<D>(type, qa, ra, S, cfg, A, B);
            }
            if (live) {
                riemann_contract<D>(cfg.flux, A, B, flags, S, cfg, phi);
            } else {
#pragma unroll
                for (int k = 0; k < U; k++) phi[k] = 0.0;
            }
#pragma unroll
            for (int k = 0; k < U; k++) Phis[f * U + k] = phi[k];
        }
        __syncthreads();
    }

    // ---- phase 3: gather, explicit Euler, residual ---------------------------------
    double r[U];
#pragma unroll
    for (int k = 0; k < U; k++) r[k] = 0.0;
    bool bad = false;
    {
        const uint16_t* slots = ta.slots + (size_t)nslot * d.cell_off;
        for (int lc = tid; lc < n_own; lc += NT) {
            double acc[U];
#pragma unroll
            for (int k = 0; k < U; k++) acc[k] = 0.0;
            for (int j = 0; j < nslot; j++) {
                const uint32_t v = slots[(size_t)j * ncgp + lc];
                if (v == 0xFFFFu) continue;
                const int lf = v >> 1;
                const double sg = (v & 1) ? -1.0 : 1.0;
#pragma unroll
                for (int k = 0; k < U; k++) acc[k] += sg * Phis[lf * U + k];
            }
            const double s = dt / ta.cvol[d.cell_off + lc];
#pragma unroll
            for (int k = 0; k < U; k++) {
                const double qo = Qs[lc * U + k];
                const double qn = qo - s * acc[k];
                Qout[lc * U + k] = qn;
                const double x = fabs(qn - qo) / qo;  // Time.cpp:72
                r[k] = fmax(r[k], (x > 0.0) ? x : 0.0);
                bad |= (qn != qn);
            }
        }
    }
    __shared__ double sm[U][NT / 32];
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < U; k++) {
        const double m = warp_max_d(r[k]);
        if (lane == 0) sm[k][wid] = m;
    }
    fence_async_smem();  // Qout (generic-proxy writes) -> visible to the bulk store
    const bool anybad = __syncthreads_or(bad);
    if (tid == 0) {
        const int even = n_own & ~1;
        if (even) bulk_s2g(Qnew + (size_t)d.cb * U, smem_u32(Qout), (uint32_t)even * U * 8u);
        if (n_own & 1)
            for (int k = 0; k < U; k++) Qnew[(size_t)(d.cb + even) * U + k] = Qout[even * U + k];
        bulk_commit_wait_read();
    }
    if (tid >= 32 && tid < 32 + U) {
        const int k = tid - 32;
        double m = 0.0;
        for (int w = 0; w < NT / 32; w++) m = fmax(m, sm[k][w]);
        if (m > 0.0) atomicMax(&resid[k], (unsigned long long)__double_as_longlong(m));
    }
    if (anybad && tid == 64) atomicOr(nanflag, 1);
}

}  // namespace mst
