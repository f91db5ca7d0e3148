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
// The Green-Gauss gradient is never formed.  rec(c,f) = Q_c + G_c.(fc_f - cc_c) with
// G_c = (1/V) sum_j Qf_j (x) Sout_j and Qf_j = eta Q[c0] + (1-eta) Q[c1] is LINEAR in
// the states of c and its face neighbours, with weights that depend on geometry only:
//     rec(c,f) = b0 Q_c + sum_j bj Q_nb(j).
// The weights are computed once at upload (csrc/tiles.cpp), so the second-order
// reconstruction of a face side is a (1 + faces-per-cell)-point weighted sum of
// cell states -- no gradient phase, no gradient storage, one barrier less.
//
// Data flow inside a CTA:
//   phase 0  TMA bulk copy of the owned block of Q (mbarrier) + 40-byte row
//            gathers of the ring cells -> Qs (shared)
//   phase 2  thread per flux face: weights / stencil ids / area vector / flags
//            streamed from the tile packet (coalesced, read once), states
//            gathered from Qs, boundary ghost, Roe / AUSM+ contracted with the
//            area vector -> Phis (shared, SoA)
//   phase 3  thread per owned cell: gather of its faces' Phis in the reference's
//            face order, Euler update in place, residual; TMA bulk store of Q
#pragma once
#include <cstdint>

#include "physics.cuh"
#include "tile_layout.h"
#include "tiles.h"

namespace mst {

struct TileArrays {
    const TileDesc* desc;
    const int32_t* ring;           // ring-cell ids, ring_stride entries per tile in desc[] order (tile t: ring + t * ring_stride)
    const unsigned char* packets;
    int32_t ring_stride;
    // streamed step (mstgpu_step_host): the launch covers entries [tile_base, tile_base + grid) of this list of
    // tile indices instead of the tiles themselves (tiles sorted by the host chunk that completes their input);
    // nullptr = the tiles in desc[] order
    const int32_t* order;
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
// pull a byte range into L2 ahead of its use (TMA prefetch, no destination)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// max over the warp of NON-NEGATIVE doubles: their order is the order of their
// bit patterns as unsigned integers, so two 32-bit REDUX do it (no shuffles)
__device__ __forceinline__ double warp_max_nonneg(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}

// resident CTAs per SM the register allocation is sized for: 2 x 256 threads or 3 x 128 threads
// (a cap of 80 registers for a third 256-thread CTA was measured slower: spills cost more than
// the extra warps gain)
#ifndef MST_TILE_MINB
#define MST_TILE_MINB(NT) ((NT) >= 256 ? 2 : 3)
#endif

// Experimental variants of the default instantiation (MSTGPU_TILE_VAR, bit mask; 0 = the measured default):
//   1  prefetch AHEAD: thread 0 pulls the packet and the owned state block of the tile this CTA slot will run
//      NEXT (persistent: its own next tile; otherwise the tile one wave of CTAs later) into L2, so that the
//      packet stream of phase 2 finds its lines in L2 from the first face on
//   2  persistent CTAs: one launch of (resident CTAs) blocks per tile class, each looping over tiles
//   4  ring rows by cp.async (LDGSTS): no register round trip, every row of a thread in flight at once
//   8 / 16  (2-D first order on triangles only) register allocation sized for 4 / 3 resident CTAs per SM
//           instead of 2 -- the DEFAULT there (4 for AUSM+, 3 for Roe: 22 % / 18 % faster); 32 forces the plain one
// All variants run the same arithmetic in the same order: results are bit-identical to variant 0.
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

// the packet words of one flux face held in registers ahead of their use (experiments VAR & 256 / 1024)
template <int D, int NS>
struct FaceWords {
    uint32_t mt, id[NS > 1 ? NS - 1 : 1];
    double S[D], w[NS > 1 ? 2 * (NS - 1) : 1];
    __device__ __forceinline__ void load(const uint32_t* __restrict__ fmeta_g, const double* __restrict__ fSd_g,
                                         const uint32_t* __restrict__ idx_g, const double* __restrict__ w_g, int nFBp, int f) {
        mt = fmeta_g[f];
#pragma unroll
        for (int dd = 0; dd < D; dd++) S[dd] = fSd_g[dd * nFBp + f];
#pragma unroll
        for (int m = 0; m < NS - 1; m++) {
            id[m] = idx_g[m * nFBp + f];
            w[m] = w_g[m * nFBp + f];
            w[NS - 1 + m] = w_g[(NS - 1 + m) * nFBp + f];
        }
    }
};

template <int D, int ORDER, int NT, int NS, bool LIM, bool VISC, int VAR>
__device__ __forceinline__ void step_tile(const TileArrays& ta, const int tile, const int pf_tile, const bool first, const bool pf_self,
                                          const uint32_t parity, int want_resid, const DevCfg& cfg, double dt_val,
                                          const double* __restrict__ dt_dev, const double* __restrict__ Qold,
                                          double* __restrict__ Qnew, unsigned long long* __restrict__ resid,
                                          int* __restrict__ nanflag) {
    constexpr int U = D + 2;
    constexpr int nslot = NS - 1;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    // Phase 0 is a chain of dependent global-memory latencies (descriptor -> ring ids -> ring rows, ~0.7 us
    // each under load).  The ring ids sit at a FIXED stride per tile, so their loads are issued together with
    // the descriptor's, before anything is known about the tile, and all rows of a thread are in flight at
    // once afterwards: two latencies instead of 1 + 2 per loop trip.
    constexpr int RB = NT >= 256 ? 1024 / NT : 4;  // ring rows per thread covered by the batch (the rest loop)
    const int32_t* __restrict__ ring_t = ta.ring + (size_t)tile * ta.ring_stride;
    int rid[RB];
#pragma unroll
    for (int j = 0; j < RB; j++) rid[j] = (tid + j * NT < ta.ring_stride) ? ring_t[tid + j * NT] : 0;
    const TileDesc d = ta.desc[tile];
    const int n_own = d.n_own, n_ring = d.n_r1 + d.n_r2;
    const int nFB = d.nFB;
    constexpr bool STG = ORDER == 2 && (VAR & 128) != 0;
    constexpr int EXT = (LIM ? 1 : 0) | (VISC ? 2 : 0) | (STG ? (4 | (NT << 8)) : 0);
    const TileLayout L = tile_layout(D, ORDER, nslot, d.n_own, d.n_r1, d.n_r2, d.nFB, EXT);
    const double dt = dt_dev ? *dt_dev : dt_val;  // device-resident dt: CFL stepping (extension)
    const int nFBp = L.nFBp, ncp = L.ncp;
    const unsigned char* pk = ta.packets + d.pk_off;
    const double* __restrict__ w_g = reinterpret_cast<const double*>(pk + L.w);
    const uint32_t* __restrict__ idx_g = reinterpret_cast<const uint32_t*>(pk + L.idx);
    const double* __restrict__ fSd_g = reinterpret_cast<const double*>(pk + L.fSd);
    const uint32_t* __restrict__ fmeta_g = reinterpret_cast<const uint32_t*>(pk + L.fmeta);
    double* Qs = reinterpret_cast<double*>(smem + L.Qs);
    double* Phis = reinterpret_cast<double*>(smem + L.Phis);  // [k][nFBp]
    const uint32_t bar = smem_u32(smem + L.mbar);

    // ---- phase 0: stage the states of the tile and its rings ------------------------
    if (first) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
    }
    if (tid == 0) {
        // bulk copies move multiples of 16 bytes = an even number of 8*U-byte rows.  An odd last row is copied by
        // hand below: rounding the bulk copy UP would write row n_own of Qs, which is the first ring row, and
        // race with the ring gather (asynchronous proxy against generic stores, no order between them).
        const uint32_t qbytes = (uint32_t)(n_own & ~1) * U * 8u;
        // slots + cvol are adjacent in the packet: one more bulk copy brings the phase-3
        // operands into shared memory without holding registers across phase 2
        const uint32_t cbytes = L.pk_bytes - L.slots;
        mbar_expect_tx(bar, qbytes + cbytes);
        if (qbytes) bulk_g2s(smem_u32(Qs), Qold + (size_t)d.cb * U, qbytes, bar);
        bulk_g2s(smem_u32(smem + L.cells_s), pk + L.slots, cbytes, bar);
        // the packet is streamed from HBM by phase 2 / 3: start moving it into L2 now,
        // while the states are being staged
        if (pf_self) bulk_prefetch_l2(pk, L.pk_bytes);
        if ((VAR & 1) && pf_tile >= 0) {
            const TileDesc dn = ta.desc[pf_tile];
            const TileLayout Ln = tile_layout(D, ORDER, nslot, dn.n_own, dn.n_r1, dn.n_r2, dn.nFB, EXT);
            bulk_prefetch_l2(ta.packets + dn.pk_off, Ln.pk_bytes);
            bulk_prefetch_l2(Qold + (size_t)dn.cb * U, (uint32_t)((dn.n_own + 1) & ~1) * U * 8u);
        }
    }
    // Staged packet stream (VAR & 128): the weights and stencil ids of a thread's NEXT face are copied into
    // its own shared-memory slots by cp.async while it works on the current one -- the bytes in flight live in
    // shared memory instead of registers, and the first use of a packet word never waits for L2 / HBM.
    // Slots are private to the thread (no barrier); [word][thread] layout, conflict-free.
    double* stg_w = reinterpret_cast<double*>(smem + L.stage);
    uint32_t* stg_i = reinterpret_cast<uint32_t*>(smem + L.stage + 2u * (NS - 1) * NT * 8u);
    auto stage_issue = [&](int f) {
#pragma unroll
        for (int m = 0; m < 2 * (NS - 1); m++) cp_async8(smem_u32(stg_w + m * NT + tid), w_g + (size_t)m * nFBp + f);
#pragma unroll
        for (int m = 0; m < NS - 1; m++) cp_async4(smem_u32(stg_i + m * NT + tid), idx_g + (size_t)m * nFBp + f);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (STG && tid < nFB) stage_issue(tid);
    if ((n_own & 1) && tid >= NT - U) {  // the odd last owned row
        const int k = tid - (NT - U);
        Qs[(n_own - 1) * U + k] = Qold[(size_t)(d.cb + n_own - 1) * U + k];
    }
    // ring cells: one thread per cell, U independent loads of a contiguous 8*U-byte row.
    // (Batching several cells per thread -- ids first, then rows -- was measured: it helps the
    // 128-thread variant but costs registers and was 1 % slower for the default 256-thread one.)
    if (VAR & 4) {
        for (int r = tid; r < n_ring; r += NT) {
            const int g = ring_t[r];
            const double* src = Qold + (size_t)g * U;
            const uint32_t dst = smem_u32(Qs + (n_own + r) * U);
#pragma unroll
            for (int k = 0; k < U; k++) cp_async8(dst + 8u * k, src + k);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        double q[RB][U];
#pragma unroll
        for (int j = 0; j < RB; j++)
            if (tid + j * NT < n_ring) {
#pragma unroll
                for (int k = 0; k < U; k++) q[j][k] = Qold[(size_t)rid[j] * U + k];
            }
#pragma unroll
        for (int j = 0; j < RB; j++)
            if (tid + j * NT < n_ring) {
#pragma unroll
                for (int k = 0; k < U; k++) Qs[(n_own + tid + j * NT) * U + k] = q[j][k];
            }
        for (int r = tid + RB * NT; r < n_ring; r += NT) {  // tiles with more than 1024 ring cells
            const int g = ring_t[r];
            double q1[U];
#pragma unroll
            for (int k = 0; k < U; k++) q1[k] = Qold[(size_t)g * U + k];
#pragma unroll
            for (int k = 0; k < U; k++) Qs[(n_own + r) * U + k] = q1[k];
        }
    }
    // VAR & 256 (experiment): the packet words of a thread's FIRST flux face are requested before the wait for the
    // staged states -- they do not depend on them, and they are the ones most likely to miss L2 (the bulk
    // prefetch of the packet was issued a moment ago): their HBM latency overlaps the ring rows' instead of
    // following it.  VAR & 1024: the words of the NEXT face are requested as soon as the reconstruction of the
    // current one is done (software pipeline through the registers the reconstruction just released).
    constexpr bool PEEL = ORDER == 2 && (VAR & (256 | 1024)) != 0 && !STG;
    constexpr bool ROT = ORDER == 2 && (VAR & 1024) != 0 && !STG;
    FaceWords<D, NS> pw;
    if constexpr (PEEL) { if (tid < nFB) pw.load(fmeta_g, fSd_g, idx_g, w_g, nFBp, tid); }
    mbar_wait(bar, parity);
    __syncthreads();

    // ---- phase 1 (limiter extension only): limiter value of every cell whose reconstruction the
    // tile evaluates (own + ring 1).  Thread per cell: min / max over the cell and its face
    // neighbours, slope at each of its face centres as a fixed-weight sum over the same stencil.
    double* philim = reinterpret_cast<double*>(smem + L.philim);  // [k][nCLp]
    if (LIM) {
        const int nCL = n_own + d.n_r1, nCLp = L.nCLp;
        const double* __restrict__ lw_g = reinterpret_cast<const double*>(pk + L.lw);
        const uint16_t* __restrict__ lid_g = reinterpret_cast<const uint16_t*>(pk + L.lid);
        const double* __restrict__ le2_g = reinterpret_cast<const double*>(pk + L.le2);
        for (int i = tid; i < nCL; i += NT) {
            // the weights stay in registers, the variables go by one at a time (register budget)
            double wj[NS - 1][NS];
            int cid[NS];
            cid[0] = i;
#pragma unroll
            for (int m = 1; m < NS; m++) cid[m] = lid_g[(m - 1) * nCLp + i];
#pragma unroll
            for (int j = 0; j < nslot; j++)
#pragma unroll
                for (int m = 0; m < NS; m++) wj[j][m] = lw_g[(j * NS + m) * nCLp + i];
            const double e2 = le2_g[i];
#pragma unroll
            for (int k = 0; k < U; k++) {
                double q[NS];
#pragma unroll
                for (int m = 0; m < NS; m++) q[m] = Qs[cid[m] * U + k];
                double qmax = q[0], qmin = q[0];
#pragma unroll
                for (int m = 1; m < NS; m++) { qmax = fmax(qmax, q[m]); qmin = fmin(qmin, q[m]); }
                const double dmax = qmax - q[0], dmin = qmin - q[0];
                LimiterAcc acc;
                acc.init(cfg.limiter);
#pragma unroll
                for (int j = 0; j < nslot; j++) {
                    double dl = wj[j][0] * q[0];
#pragma unroll
                    for (int m = 1; m < NS; m++) dl += wj[j][m] * q[m];
                    acc.add(cfg.limiter, dl, dmax, dmin, e2);
                }
                philim[k * nCLp + i] = acc.phi(cfg.limiter, dmax, dmin);
            }
        }
        __syncthreads();
    }

    // ---- phase 1v (viscous extension only): Green-Gauss gradient of the face primitives (u_i, T) of every
    // cell next to a flux face (own + ring 1), as k_gradient forms it: face state = eta-weighted mean of the
    // two cell states, primitives of THAT state, summed with the outward area vectors (pre-divided by V).
    constexpr int P = D + 1;
    double* Gps = reinterpret_cast<double*>(smem + L.Gps);  // [(k*D+d)][nCLp]
    if (VISC) {
        const int nCL = n_own + d.n_r1, nCLp = L.nCLp;
        const double* __restrict__ vw_g = reinterpret_cast<const double*>(pk + L.vw);
        const uint16_t* __restrict__ lid_g = reinterpret_cast<const uint16_t*>(pk + L.lid);
        constexpr int W = 2 + D;
        for (int i = tid; i < nCL; i += NT) {
            double q0[U], t[P][D];
#pragma unroll
            for (int k = 0; k < U; k++) q0[k] = Qs[i * U + k];
#pragma unroll
            for (int k = 0; k < P; k++)
#pragma unroll
                for (int dd = 0; dd < D; dd++) t[k][dd] = 0.0;
#pragma unroll
            for (int j = 0; j < nslot; j++) {
                const int c = lid_g[j * nCLp + i];
                const double e0 = vw_g[(j * W + 0) * nCLp + i], e1 = vw_g[(j * W + 1) * nCLp + i];
                double qf[U], pr[P], m2 = 0.0;
#pragma unroll
                for (int k = 0; k < U; k++) qf[k] = e0 * q0[k] + e1 * Qs[c * U + k];
                // one FP64 division per face state (the split path divides six times: ~20 issue slots each)
                const double rf = 1.0 / qf[0];
#pragma unroll
                for (int a = 0; a < D; a++) { pr[a] = qf[a + 1] * rf; m2 += qf[a + 1] * qf[a + 1]; }
                pr[D] = (qf[U - 1] - 0.5 * m2 * rf) * rf * cfg.inv_cv;  // T, FUNCTION.cpp:8-11
#pragma unroll
                for (int dd = 0; dd < D; dd++) {
                    const double sv = vw_g[(j * W + 2 + dd) * nCLp + i];
#pragma unroll
                    for (int k = 0; k < P; k++) t[k][dd] += pr[k] * sv;
                }
            }
#pragma unroll
            for (int k = 0; k < P; k++)
#pragma unroll
                for (int dd = 0; dd < D; dd++) Gps[(k * D + dd) * nCLp + i] = t[k][dd];
        }
        __syncthreads();
    }

    // ---- phase 2: reconstruction (fixed stencil) + flux on every face with an owned cell ----
    for (int f = tid; f < nFB; f += NT) {
        if constexpr (PEEL && !ROT) { if (f != tid) pw.load(fmeta_g, fSd_g, idx_g, w_g, nFBp, f); }
        if constexpr ((VAR & 512) != 0) if (f + NT < nFB && (tid & 15) == 0) {
            // experiment: the next trip's packet lines into L1 (one request per 128-byte line)
            const int fn = f + NT;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(fmeta_g + fn));
#pragma unroll
            for (int dd = 0; dd < D; dd++) asm volatile("prefetch.global.L1 [%0];" ::"l"(fSd_g + dd * nFBp + fn));
#pragma unroll
            for (int m = 0; m < NS - 1; m++) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(idx_g + m * nFBp + fn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(w_g + m * nFBp + fn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(w_g + (NS - 1 + m) * nFBp + fn));
            }
        }
        uint32_t mt;
        double S[D];
        if constexpr (PEEL) {
            mt = pw.mt;
#pragma unroll
            for (int dd = 0; dd < D; dd++) S[dd] = pw.S[dd];
        } else {
            mt = fmeta_g[f];
#pragma unroll
            for (int dd = 0; dd < D; dd++) S[dd] = fSd_g[dd * nFBp + f];
        }
        const int type = mt & 0xff;
        const uint32_t flags = mt >> 8;
        double A[U], B[U], phi[U];
        double visc[VISC ? U : 1];
        bool live = true;
        if (ORDER == 2) {
            uint32_t id[NS - 1];
            double wa[NS], wb[NS];
            if (STG) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
                for (int m = 0; m < NS - 1; m++) {
                    id[m] = stg_i[m * NT + tid];
                    wa[m + 1] = stg_w[m * NT + tid];
                    wb[m + 1] = stg_w[(NS - 1 + m) * NT + tid];
                }
                if (f + NT < nFB) stage_issue(f + NT);  // the slots were just read by their only reader
            } else if constexpr (PEEL) {
#pragma unroll
                for (int m = 0; m < NS - 1; m++) {
                    id[m] = pw.id[m];
                    wa[m + 1] = pw.w[m];
                    wb[m + 1] = pw.w[NS - 1 + m];
                }
            } else {
#pragma unroll
                for (int m = 0; m < NS - 1; m++) {
                    id[m] = idx_g[m * nFBp + f];
                    wa[m + 1] = w_g[m * nFBp + f];
                    wb[m + 1] = w_g[(NS - 1 + m) * nFBp + f];
                }
            }
            // own-cell weight: a closed cell reproduces constants (checked when the packets were built)
            double sa = wa[1], sb = wb[1];
#pragma unroll
            for (int m = 2; m < NS; m++) { sa += wa[m]; sb += wb[m]; }
            wa[0] = 1.0 - sa;
            wb[0] = 1.0 - sb;
            const int la = id[0] & 0xFFFFu, lb = id[0] >> 16;
            const bool interior = lb != 0xFFFF;
            const int lbx = interior ? lb : la;  // boundary face: the same row twice (one address select, no predicated loads)
            // the two cells of the face serve both sides: own cell of one, first neighbour of the other
            double qa[U], qb[U];
#pragma unroll
            for (int k = 0; k < U; k++) {
                qa[k] = Qs[la * U + k];
                qb[k] = Qs[lbx * U + k];
                A[k] = wa[0] * qa[k] + wa[1] * qb[k];
                B[k] = wb[0] * qb[k] + wb[1] * qa[k];
            }
#pragma unroll
            for (int m = 2; m < NS; m++) {
                const int ca = id[m - 1] & 0xFFFFu;
#pragma unroll
                for (int k = 0; k < U; k++) A[k] += wa[m] * Qs[ca * U + k];
            }
            if (interior) {
#pragma unroll
                for (int m = 2; m < NS; m++) {
                    const int cb = id[m - 1] >> 16;
#pragma unroll
                    for (int k = 0; k < U; k++) B[k] += wb[m] * Qs[cb * U + k];
                }
                if (LIM) {
#pragma unroll
                    for (int k = 0; k < U; k++) {
                        A[k] = qa[k] + philim[k * L.nCLp + la] * (A[k] - qa[k]);
                        B[k] = qb[k] + philim[k * L.nCLp + lb] * (B[k] - qb[k]);
                    }
                }
            } else {
                if (LIM) {
#pragma unroll
                    for (int k = 0; k < U; k++) A[k] = qa[k] + philim[k * L.nCLp + la] * (A[k] - qa[k]);
                }
                double ra[U];
#pragma unroll
                for (int k = 0; k < U; k++) ra[k] = A[k];
                live = boundary_states<D>(type, qa, ra, S, cfg, A, B);
            }
            if (VISC) {
                // laminar viscous flux (corrected formulation, see k_flux): tau = mu (grad u + grad u^T) +
                // lambda div(u) I, energy flux u.tau + k grad T; face values by the eta interpolation.
                // Kept in phi[] and combined with the convective flux below.
                const double e = reinterpret_cast<const double*>(pk + L.feta)[f];
                const int nCLp = L.nCLp;
                double qf[U], gf[P][D];
                if (interior) {
#pragma unroll
                    for (int k = 0; k < U; k++) qf[k] = e * qa[k] + (1.0 - e) * qb[k];
#pragma unroll
                    for (int k = 0; k < P; k++)
#pragma unroll
                        for (int dd = 0; dd < D; dd++)
                            gf[k][dd] = e * Gps[(k * D + dd) * nCLp + la] + (1.0 - e) * Gps[(k * D + dd) * nCLp + lb];
                } else {
#pragma unroll
                    for (int k = 0; k < U; k++) qf[k] = qa[k];
#pragma unroll
                    for (int k = 0; k < P; k++)
#pragma unroll
                        for (int dd = 0; dd < D; dd++) gf[k][dd] = Gps[(k * D + dd) * nCLp + la];
                }
                double uf[D], div = 0.0;
                const double rqf = 1.0 / qf[0];
#pragma unroll
                for (int i = 0; i < D; i++) { uf[i] = qf[i + 1] * rqf; div += gf[i][i]; }
                double en = 0.0, fv[D];
#pragma unroll
                for (int i = 0; i < D; i++) fv[i] = 0.0;
#pragma unroll
                for (int j = 0; j < D; j++) {
                    double w = 0.0;
#pragma unroll
                    for (int i = 0; i < D; i++) {
                        double tij = cfg.mu * (gf[i][j] + gf[j][i]);
                        if (i == j) tij += cfg.lambda * div;
                        fv[i] += tij * S[j];
                        w += uf[i] * tij;
                    }
                    en += (w + cfg.kappa * gf[D][j]) * S[j];
                }
                visc[0] = 0.0;
#pragma unroll
                for (int i = 0; i < D; i++) visc[i + 1] = fv[i];
                visc[U - 1] = en;
            }
        } else {
            const uint32_t ab = idx_g[f];
            const int la = ab & 0xFFFFu, lb = ab >> 16;
            double qa[U];
#pragma unroll
            for (int k = 0; k < U; k++) qa[k] = Qs[la * U + k];
            if (lb != 0xFFFF) {
#pragma unroll
                for (int k = 0; k < U; k++) { A[k] = qa[k]; B[k] = Qs[lb * U + k]; }
            } else {
                live = boundary_states<D>(type, qa, qa, S, cfg, A, B);
            }
        }
        if constexpr (ROT) { if (f + NT < nFB) pw.load(fmeta_g, fSd_g, idx_g, w_g, nFBp, f + NT); }  // S of the current face was copied out above
        if (live) {
            riemann_contract<D>(cfg.flux, A, B, flags, S, cfg, phi);
        } else {
#pragma unroll
            for (int k = 0; k < U; k++) phi[k] = 0.0;
        }
        if (VISC) {
#pragma unroll
            for (int k = 0; k < U; k++) phi[k] -= visc[k];
        }
#pragma unroll
        for (int k = 0; k < U; k++) Phis[k * nFBp + f] = phi[k];
    }
    __syncthreads();

    // ---- phase 3: gather, explicit Euler (in place in Qs), residual -------------------
    double r[U];
#pragma unroll
    for (int k = 0; k < U; k++) r[k] = 0.0;
    bool bad = false;
    const uint16_t* slots_s = reinterpret_cast<const uint16_t*>(smem + L.cells_s);
    const double* cvol_s = reinterpret_cast<const double*>(smem + L.cells_s + (L.cvol - L.slots));
    for (int lc = tid; lc < n_own; lc += NT) {
        double acc[U];
#pragma unroll
        for (int k = 0; k < U; k++) acc[k] = 0.0;
#pragma unroll
        for (int j = 0; j < nslot; j++) {
            const uint32_t v = slots_s[j * ncp + lc];
            if (v == 0xFFFFu) continue;
            const int lf = v >> 1;
            const double sg = (v & 1) ? -1.0 : 1.0;
#pragma unroll
            for (int k = 0; k < U; k++) acc[k] += sg * Phis[k * nFBp + lf];
        }
        const double s = dt * cvol_s[lc];  // DT / V (RhoSolver.cpp:64); the packet stores 1/V
#pragma unroll
        for (int k = 0; k < U; k++) {
            const double qo = Qs[lc * U + k];
            // bit 1 of want_resid: residual-vector mode (implicit step): -R_i instead of the update
            const double qn = (want_resid & 2) ? -acc[k] : qo - s * acc[k];
            Qs[lc * U + k] = qn;
            if (want_resid & 1) {  // only the last step of a multi-step call can be observed (mstgpu_residual_linf)
                const double x = fabs(qn - qo) * __drcp_rn(qo);  // Time.cpp:72 (|d|/q; reporting only, 1 ulp)
                r[k] = fmax(r[k], (x > 0.0) ? x : 0.0);
            }
            bad |= (qn != qn);
        }
    }
    __shared__ double sm[U][NT / 32];
    const int lane = tid & 31, wid = tid >> 5;
    if (want_resid & 1) {
#pragma unroll
        for (int k = 0; k < U; k++) {
            const double m = warp_max_nonneg(r[k]);
            if (lane == 0) sm[k][wid] = m;
        }
    }
    fence_async_smem();  // generic-proxy writes of Qs -> visible to the bulk store
    const bool anybad = __syncthreads_or(bad);
    if (tid == 0) {
        const int even = n_own & ~1;
        if (even) bulk_s2g(Qnew + (size_t)d.cb * U, smem_u32(Qs), (uint32_t)even * U * 8u);
        if (n_own & 1)
            for (int k = 0; k < U; k++) Qnew[(size_t)(d.cb + even) * U + k] = Qs[even * U + k];
        bulk_commit_wait_read();
    }
    if ((want_resid & 1) && tid >= 32 && tid < 32 + U) {
        const int k = tid - 32;
        double m = 0.0;
        for (int w = 0; w < NT / 32; w++) m = fmax(m, sm[k][w]);
        if (m > 0.0) atomicMax(&resid[k], (unsigned long long)__double_as_longlong(m));
    }
    if (anybad && tid == 64) atomicOr(nanflag, 1);
}

// n_class = tiles in this launch's class (persistent loop bound, prefetch bound); var_arg = prefetch distance in
// tiles for the non-persistent prefetch-ahead variant (one wave of resident CTAs)
template <int D, int ORDER, int NT, int NS, bool LIM = false, bool VISC = false, int VAR = 0>
__global__ void __launch_bounds__(NT, (VAR & 64) ? 5 : (VAR & 8) ? 4 : (VAR & 16) ? 3 : MST_TILE_MINB(NT)) k_step_tiles(TileArrays ta, int tile_base, int n_class, int var_arg,
                                                   int want_resid, DevCfg cfg, double dt_val,
                                                   const double* __restrict__ dt_dev,
                                                   const double* __restrict__ Qold,
                                                   double* __restrict__ Qnew,
                                                   unsigned long long* __restrict__ resid,
                                                   int* __restrict__ nanflag) {
    if (VAR & 2) {
        uint32_t it = 0;
        for (int t = blockIdx.x; t < n_class; t += gridDim.x, it++) {
            const int tn = t + (int)gridDim.x;
            step_tile<D, ORDER, NT, NS, LIM, VISC, VAR>(ta, tile_base + t, ((VAR & 1) && tn < n_class) ? tile_base + tn : -1, it == 0, !(VAR & 1) || it == 0, it & 1u,
                                                        want_resid, cfg, dt_val, dt_dev, Qold, Qnew, resid, nanflag);
            __syncthreads();  // the bulk store has read Qs (thread 0 waited for it): the next tile may overwrite it
        }
    } else {
        const int tn = (int)blockIdx.x + var_arg;
        const int tile = ta.order ? ta.order[tile_base + (int)blockIdx.x] : tile_base + (int)blockIdx.x;
        step_tile<D, ORDER, NT, NS, LIM, VISC, VAR>(ta, tile, ((VAR & 1) && tn < n_class) ? tile_base + tn : -1,
                                                    true, !(VAR & 1) || (int)blockIdx.x < var_arg, 0u, want_resid, cfg, dt_val, dt_dev, Qold,
                                                    Qnew, resid, nanflag);
    }
}

}  // namespace mst
