// lusgs.cu -- LU-SGS sweeps of the reference's lusolver on the GPU (C ABI in
// include/mstgpu.h, mstgpu_lusgs_*).  Reference (R = /root/reference/MST-CFD):
//   scalar  SparseSolverNUM::solveILUSGS    R/lusolver/SparseSolverNUM.cpp:144-212
//   block   SparseSolver<MT,VCT>::solveILU  R/lusolver/SparseSolver.cpp:54-104
//
// The reference sweeps are strictly sequential (forward i = 0..n-1, backward
// i = n-1..0, column-stored scatter).  Here the same recurrences run LEVEL BY
// LEVEL: row r of the forward sweep only needs the final values of the rows
// c < r it is coupled to, so all rows whose dependencies are complete form one
// level and are processed by one launch.  Each row subtracts its terms in the
// reference's order (ascending column forward, descending backward), so the
// result equals the sequential sweep on the SAME matrix ordering -- no
// reordering is imposed.  The number of levels is what the ordering makes it:
// a colour-ordered matrix (mstgpu_lusgs_color_order) has one level per colour,
// a lexicographic one a level per wavefront.
//
// Internal numbering = SWEEP POSITION.  Everything the solver owns (the L / U / D index lists, the scaled blocks
// LD / UD, D, D^-1 and the work vectors) is indexed by the position p of a row in the sweep, so the rows of one
// level -- one colour -- are a contiguous range and a level's launch streams its blocks and vectors front to
// back.  Only the caller's arrays keep the caller's numbering: val is reached through the entry positions,
// b and x through rmap[p] = row of sweep position p (ghost columns >= n keep their index).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mstgpu.h"

namespace {
thread_local std::string g_lusgs_error;
}

struct mstgpu_lusgs {
    int n = 0, B = 1, device = 0;
    cudaStream_t stream = nullptr;
    int nnz = 0, nL = 0, nU = 0, nG = 0, ncols = 0;
    // strictly-lower / strictly-upper CSR by row; pos = index of the entry in the caller's val array
    int *Lptr = nullptr, *Lcol = nullptr, *Lpos = nullptr, *Uptr = nullptr, *Ucol = nullptr, *Upos = nullptr;
    int *Dptr = nullptr, *Dpos = nullptr;  // diagonal entries per row (summed)
    int *Gptr = nullptr, *Gcol = nullptr, *Gpos = nullptr;  // entries in ghost columns (>= n): lagged, moved to the right-hand side
    int* rmap = nullptr;  // [n] sweep position -> row of the caller's b / x
    int* rinv = nullptr;  // [n] row of the caller -> sweep position
    int setup_mode = 2;   // MSTGPU_LUSGS_SETUP: 2 = k_diag_reg + k_scale_rows (default), 1 = k_diag_grp + k_scale_rows, 0 = k_diag + k_scale (round 1)
    double* beff = nullptr;
    std::vector<int> fptr, bptr;           // level pointers (host)
    int *frows = nullptr, *brows = nullptr;
    double *val = nullptr, *D = nullptr, *Dinv = nullptr, *LD = nullptr, *UD = nullptr;
    double *b = nullptr, *x = nullptr, *rhs = nullptr, *rhs1 = nullptr, *ux = nullptr;
    double* s = nullptr;  // mode 1: U x of the next iteration, a by-product of the backward sweep
    int mode = 2;         // 0 = the reference's four passes per iteration, 1 = fused, 2 = lean (default; see solve_core)
    bool s_valid = false; // h->s holds U x of the CURRENT x (left behind by the last backward sweep of modes 1 / 2)
    bool x0_zero = false; // hint of the caller for the next solve: the start vector is zero (U x = 0, no pass over U)
    unsigned long long* res = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    std::string err;
};

#define LCK(call)                                                                \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            g_lusgs_error = std::string(#call) + ": " + cudaGetErrorString(e_);  \
            if (h) h->err = g_lusgs_error;                                       \
            return MSTGPU_ERR_CUDA;                                              \
        }                                                                        \
    } while (0)

namespace {

template <int B>
__device__ __forceinline__ void d_matvec(const double* m, const double* v, double* out) {
#pragma unroll
    for (int i = 0; i < B; i++) {
        double s = m[i * B] * v[0];
#pragma unroll
        for (int k = 1; k < B; k++) s += m[i * B + k] * v[k];
        out[i] = s;
    }
}

template <int B>
__device__ void d_inverse(const double* m, double* inv) {
    if (B == 1) { inv[0] = 1.0 / m[0]; return; }
    double w[B][2 * B];
    for (int i = 0; i < B; i++)
        for (int j = 0; j < B; j++) { w[i][j] = m[i * B + j]; w[i][B + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < B; c++) {
        int p = c;
        for (int r = c + 1; r < B; r++) if (fabs(w[r][c]) > fabs(w[p][c])) p = r;
        if (p != c) for (int j = 0; j < 2 * B; j++) { double t = w[c][j]; w[c][j] = w[p][j]; w[p][j] = t; }
        const double ip = 1.0 / w[c][c];
        for (int j = 0; j < 2 * B; j++) w[c][j] *= ip;
        for (int r = 0; r < B; r++) {
            if (r == c) continue;
            const double f = w[r][c];
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * B; j++) w[r][j] -= f * w[c][j];
        }
    }
    for (int i = 0; i < B; i++) for (int j = 0; j < B; j++) inv[i * B + j] = w[i][B + j];
}

// ---- thread mapping ---------------------------------------------------------------------------------
// A block row (B scalar rows) is handled by B adjacent lanes, lane i owning scalar row i of every
// B x B block it meets: consecutive lanes read consecutive 8*B-byte block rows, so the block arrays
// (AoS, 8*B*B bytes per entry, rows in storage order) stream through fully used sectors, and the
// vectors are written B doubles per group.  A group never straddles a warp: a warp takes
// RPW = 32 / B block rows (lanes >= RPW * B idle), so the B values of a row can be exchanged with
// shuffles where one lane needs its siblings' results (M v with v computed by the group).
template <int B>
struct Grp {
    static constexpr int RPW = 32 / B;
    int row, i, base;  // index of the group's item, scalar row within the block, first lane of the group
    bool on;
    __device__ __forceinline__ Grp(int nitems) {
        const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int rl = lane / B;
        i = lane - rl * B;
        base = rl * B;
        row = warp * RPW + rl;
        on = rl < RPW && row < nitems;
    }
    static int grid(int nitems, int threads) {
        const int warps = (nitems + RPW - 1) / RPW, wpb = threads / 32;
        return (warps + wpb - 1) / wpb;
    }
    // (M v)_i where lane k of the group holds v_k; every lane of the warp must call it
    __device__ __forceinline__ double matvec_lanes(const double* m_row, double vi) const {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < B; k++) {
            const double vk = __shfl_sync(0xffffffffu, vi, base + k);
            if (on) s = (k == 0) ? m_row[0] * vk : s + m_row[k] * vk;
        }
        return s;
    }
};

// row i of a B x B block times a vector in memory: same order of operations as d_matvec
template <int B>
__device__ __forceinline__ double row_dot(const double* m_row, const double* v) {
    double s = m_row[0] * v[0];
#pragma unroll
    for (int k = 1; k < B; k++) s += m_row[k] * v[k];
    return s;
}

// D = sum of the row's diagonal entries (addD), Dinv = D^-1
template <int B>
__global__ void k_diag(int n, const int* Dptr, const int* Dpos, const double* val, double* D, double* Dinv) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double d[B * B];
    for (int q = 0; q < B * B; q++) d[q] = 0.0;
    for (int k = Dptr[r]; k < Dptr[r + 1]; k++)
        for (int q = 0; q < B * B; q++) d[q] += val[(size_t)Dpos[k] * B * B + q];
    double di[B * B];
    d_inverse<B>(d, di);
    for (int q = 0; q < B * B; q++) { D[(size_t)r * B * B + q] = d[q]; Dinv[(size_t)r * B * B + q] = di[q]; }
}

// D, D^-1 with a lane group per block row (lane i owns row i of [D | I]); the Gauss-Jordan elimination with partial
// pivoting of d_inverse, same operations in the same order, rows exchanged and broadcast by shuffles -- no local
// memory.  Rows are taken in the CALLER's order r (val streams front to back), results stored at p = rinv[r].
template <int B>
__global__ void k_diag_grp(int n, const int* __restrict__ rinv, const int* __restrict__ Dptr, const int* __restrict__ Dpos,
                           const double* __restrict__ val, double* __restrict__ D, double* __restrict__ Dinv) {
    const Grp<B> g(n);
    const int i = g.i, base = g.base;
    const int p = g.on ? rinv[g.row] : 0;
    double w[2 * B];
#pragma unroll
    for (int j = 0; j < B; j++) { w[j] = 0.0; w[B + j] = (i == j) ? 1.0 : 0.0; }
    if (g.on)
        for (int k = Dptr[p]; k < Dptr[p + 1]; k++) {
#pragma unroll
            for (int j = 0; j < B; j++) w[j] += val[(size_t)Dpos[k] * B * B + i * B + j];
        }
    if (g.on) {
#pragma unroll
        for (int j = 0; j < B; j++) D[(size_t)p * B * B + i * B + j] = w[j];
    } else {
        w[i < B ? i : 0] = 1.0;  // idle lanes run the elimination on the identity: no 0/0 in the shuffles
    }
    if (B == 1) { if (g.on) Dinv[p] = 1.0 / w[0]; return; }
#pragma unroll
    for (int c = 0; c < B; c++) {
        // pivot row: the first largest |w[r][c]| over r >= c (d_inverse)
        const double mine = (i >= c) ? fabs(w[c]) : -1.0;
        double best = __shfl_sync(0xffffffffu, mine, base + c);
        int piv = c;
#pragma unroll
        for (int r = c + 1; r < B; r++) {
            const double v = __shfl_sync(0xffffffffu, mine, base + r);
            if (v > best) { best = v; piv = r; }
        }
        // exchange rows c and piv (a no-op when piv == c), then scale row c and eliminate column c elsewhere
        double rowc[2 * B];
#pragma unroll
        for (int j = 0; j < 2 * B; j++) {
            const double a = __shfl_sync(0xffffffffu, w[j], base + c), b = __shfl_sync(0xffffffffu, w[j], base + piv);
            if (i == piv) w[j] = a;
            if (i == c) w[j] = b;   // after the line above: piv == c leaves the row as it is
            rowc[j] = b;            // the new row c, before scaling
        }
        const double ip = 1.0 / rowc[c];
#pragma unroll
        for (int j = 0; j < 2 * B; j++) rowc[j] *= ip;
        if (i == c) {
#pragma unroll
            for (int j = 0; j < 2 * B; j++) w[j] = rowc[j];
        } else {
            const double f = w[c];
            if (f != 0.0) {
#pragma unroll
                for (int j = 0; j < 2 * B; j++) w[j] -= f * rowc[j];
            }
        }
    }
    if (g.on) {
#pragma unroll
        for (int j = 0; j < B; j++) Dinv[(size_t)p * B * B + i * B + j] = w[B + j];
    }
}

// D, D^-1 with a thread per block row and the augmented matrix in REGISTERS: d_inverse's Gauss-Jordan elimination
// (same pivot rule, same operations in the same order) fully unrolled, rows exchanged by predicated swaps -- no
// local memory (k_diag keeps [D | I] in a dynamically indexed local array) and no shuffles (k_diag_grp spends
// ~250 of them per row).  Memory side: a warp fetches its 32 diagonal blocks one 8*B*B-byte row per request
// (lanes < B*B) into shared memory and writes D and D^-1 as two contiguous 32-row runs.
template <int B, int C>
__device__ __forceinline__ void d_inverse_col(double (&w)[B][2 * B]) {
    if constexpr (C < B) {
        int p = C;
        double best = fabs(w[C][C]);
#pragma unroll
        for (int r = C + 1; r < B; r++) {
            const double v = fabs(w[r][C]);
            if (v > best) { best = v; p = r; }
        }
#pragma unroll
        for (int r = C + 1; r < B; r++) {
            const bool sw = p == r;
#pragma unroll
            for (int j = 0; j < 2 * B; j++) {
                const double a = w[C][j], b = w[r][j];
                w[C][j] = sw ? b : a;
                w[r][j] = sw ? a : b;
            }
        }
        const double ip = 1.0 / w[C][C];
#pragma unroll
        for (int j = 0; j < 2 * B; j++) w[C][j] *= ip;
#pragma unroll
        for (int r = 0; r < B; r++) {
            if (r != C) {
                const double f = w[r][C];
                if (f != 0.0) {
#pragma unroll
                    for (int j = 0; j < 2 * B; j++) w[r][j] -= f * w[C][j];
                }
            }
        }
        d_inverse_col<B, C + 1>(w);
    }
}
template <int B>
__device__ __forceinline__ void d_inverse_reg(double (&w)[B][2 * B]) { d_inverse_col<B, 0>(w); }

template <int B>
__global__ void __launch_bounds__(128) k_diag_reg(int n, const int* __restrict__ Dptr, const int* __restrict__ Dpos,
                                                  const double* __restrict__ val, double* __restrict__ D, double* __restrict__ Dinv) {
    constexpr int BB = B * B;
    __shared__ double sh[4][32 * BB];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int r0 = (blockIdx.x * 4 + wib) * 32;
    if (r0 >= n) return;  // whole warp
    const int nr = min(32, n - r0);
    double* sw = sh[wib];
    // diagonal entries of "my" row; the common case is exactly one (its position is fetched ahead of the loop)
    const int k0 = lane < nr ? Dptr[r0 + lane] : 0, k1 = lane < nr ? Dptr[r0 + lane + 1] : 0;
    const int pos0 = k1 > k0 ? Dpos[k0] : -1;
#pragma unroll 8
    for (int i = 0; i < nr; i++) {
        const int a0 = __shfl_sync(0xffffffffu, k0, i), a1 = __shfl_sync(0xffffffffu, k1, i), ps = __shfl_sync(0xffffffffu, pos0, i);
        if (lane < BB) {
            double a = ps >= 0 ? val[(size_t)ps * BB + lane] : 0.0;
            for (int k = a0 + 1; k < a1; k++) a += val[(size_t)Dpos[k] * BB + lane];  // addD: several entries on the diagonal
            sw[i * BB + lane] = a;
        }
    }
    __syncwarp();
    for (int t = lane; t < nr * BB; t += 32) D[(size_t)r0 * BB + t] = sw[t];
    double w[B][2 * B];
    if (lane < nr) {
#pragma unroll
        for (int i = 0; i < B; i++)
#pragma unroll
            for (int j = 0; j < B; j++) { w[i][j] = sw[lane * BB + i * B + j]; w[i][B + j] = (i == j) ? 1.0 : 0.0; }
        if (B == 1) w[0][1] = 1.0 / w[0][0];
        else d_inverse_reg<B>(w);
    }
    __syncwarp();
    if (lane < nr) {
#pragma unroll
        for (int i = 0; i < B; i++)
#pragma unroll
            for (int j = 0; j < B; j++) sw[lane * BB + i * B + j] = w[i][B + j];
    }
    __syncwarp();
    for (int t = lane; t < nr * BB; t += 32) Dinv[(size_t)r0 * BB + t] = sw[t];
}

// LD / UD of one block row per lane group, rows in the caller's order: val streams front to back and the D^-1 blocks
// of a row's neighbours -- close in the caller's (Hilbert) order, so used again within a few rows -- stay in L2,
// instead of being fetched once per colour pass.  Same arithmetic as k_scale.
template <int B>
__global__ void k_scale_rows(int n, const int* __restrict__ rinv, const int* __restrict__ Lptr, const int* __restrict__ Lcol,
                             const int* __restrict__ Lpos, const int* __restrict__ Uptr, const int* __restrict__ Ucol,
                             const int* __restrict__ Upos, const double* __restrict__ val, const double* __restrict__ D,
                             const double* __restrict__ Dinv, double* __restrict__ LD, double* __restrict__ UD) {
    const Grp<B> g(n);
    if (!g.on) return;
    const int i = g.i, p = rinv[g.row];
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int* ptr = half ? Uptr : Lptr;
        const int* col = half ? Ucol : Lcol;
        const int* pos = half ? Upos : Lpos;
        double* XD = half ? UD : LD;
        for (int e = ptr[p]; e < ptr[p + 1]; e++) {
            const double* a = val + (size_t)pos[e] * B * B + i * B;
            if (B == 1) { XD[e] = a[0] * (1. / D[col[e]]); continue; }  // SparseSolverNUM.cpp:181: L * (1./D)
            const double* di = Dinv + (size_t)col[e] * B * B;
#pragma unroll
            for (int j = 0; j < B; j++) {
                double s = a[0] * di[j];
#pragma unroll
                for (int k = 1; k < B; k++) s += a[k] * di[k * B + j];
                XD[(size_t)e * B * B + i * B + j] = s;
            }
        }
    }
}

// XD[e] = X[e] * Dinv[col[e]]   (the reference's (L * D^-1) factor, SparseSolver.cpp:86,96); thread per (entry, row)
template <int B>
__global__ void k_scale(size_t ne, const int* col, const int* pos, const double* val, const double* D, const double* Dinv,
                        double* XD) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ne * B) return;
    const size_t e = g / B;
    const int i = (int)(g - e * B);
    const double* a = val + (size_t)pos[e] * B * B + i * B;
    if (B == 1) { XD[e] = a[0] * (1. / D[col[e]]); return; }  // SparseSolverNUM.cpp:181: L * (1./D)
    const double* di = Dinv + (size_t)col[e] * B * B;
#pragma unroll
    for (int j = 0; j < B; j++) {
        double s = a[0] * di[j];
#pragma unroll
        for (int k = 1; k < B; k++) s += a[k] * di[k * B + j];
        XD[e * B * B + i * B + j] = s;
    }
}

// ux[r] = D^-1 (sum_{c after r} U[r,c] x[c])     (SparseSolverNUM.cpp:158-166)
template <int B>
__global__ void k_ux(int n, const int* Uptr, const int* Ucol, const int* Upos, const int* rmap, const double* val, const double* D,
                     const double* Dinv, const double* x, double* ux) {
    const Grp<B> g(n);
    const int r = g.row, i = g.i;
    double acc = 0.0;
    if (g.on)
        for (int k = Uptr[r]; k < Uptr[r + 1]; k++)
            acc += row_dot<B>(val + (size_t)Upos[k] * B * B + i * B, x + (size_t)rmap[Ucol[k]] * B);
    if (B == 1) { if (g.on) ux[r] = acc / D[r]; return; }
    const double t = g.matvec_lanes(g.on ? Dinv + (size_t)r * B * B + i * B : nullptr, acc);
    if (g.on) ux[(size_t)r * B + i] = t;
}

// beff[r] = b[r] - sum_{ghost columns c} A[r,c] x[c]: couplings to rows another partition owns, with the
// x the caller supplied for them (the previous iteration's values: block Jacobi across partitions)
template <int B>
__global__ void k_ghost_rhs(int n, const int* Gptr, const int* Gcol, const int* Gpos, const int* rmap, const double* val, const double* b,
                            const double* x, double* beff) {
    const Grp<B> g(n);
    if (!g.on) return;
    const int r = g.row, i = g.i;
    double acc = b[(size_t)rmap[r] * B + i];
    for (int k = Gptr[r]; k < Gptr[r + 1]; k++)
        acc -= row_dot<B>(val + (size_t)Gpos[k] * B * B + i * B, x + (size_t)Gcol[k] * B);
    beff[(size_t)r * B + i] = acc;
}

// rhs[r] = b[r] + sum_{c before r} L[r,c] ux[c]   (SparseSolverNUM.cpp:167-175)
template <int B>
__global__ void k_rhs(int n, const int* Lptr, const int* Lcol, const int* Lpos, const int* rmap, const double* val, const double* b,
                      const double* ux, double* rhs) {
    const Grp<B> g(n);
    if (!g.on) return;
    const int r = g.row, i = g.i;
    double acc = 0.0;
    for (int k = Lptr[r]; k < Lptr[r + 1]; k++)
        acc += row_dot<B>(val + (size_t)Lpos[k] * B * B + i * B, ux + (size_t)Lcol[k] * B);
    rhs[(size_t)r * B + i] = b[(size_t)(rmap ? rmap[r] : r) * B + i] + acc;
}

// one level of a triangular sweep: v[r] -= sum_k XD[k] v[col[k]], k in sweep order (forward) or reversed (backward)
template <int B, bool FWD>
__global__ void k_sweep_level(int nrows, const int* rows, const int* ptr, const int* col, const double* XD, double* v) {
    const Grp<B> g(nrows);
    if (!g.on) return;
    const int r = rows[g.row], i = g.i;
    double acc = v[(size_t)r * B + i];
    const int k0 = ptr[r], k1 = ptr[r + 1];
    for (int kk = 0; kk < k1 - k0; kk++) {
        const int k = FWD ? k0 + kk : k1 - 1 - kk;
        acc -= row_dot<B>(XD + (size_t)k * B * B + i * B, v + (size_t)col[k] * B);
    }
    v[(size_t)r * B + i] = acc;
}

// ---- fused iteration (mode 1) ------------------------------------------------------------------------
// With XD[r,c] = X[r,c] D_c^-1 (k_scale) the reference iteration (SparseSolver.cpp:54-104) is, in RHS space,
//     s = U x;  rhs = b + LD s;  v[r] = rhs[r] - sum_{c<r} LD[r,c] v[c];  w0 = D (D^-1 v);
//     w[r] = w0[r] - sum_{c>r} UD[r,c] w[c];  x = D^-1 w.
// The backward sweep already forms sum_{c>r} UD[r,c] w[c] = sum U[r,c] x[c] = s[r] of the NEXT iteration, and
// rhs + forward sweep collapse into v[r] = b[r] + sum_{c<r} LD[r,c] (s[c] - v[c]).  The unscaled blocks are then
// read once per solve (first s) instead of twice per iteration: off-diagonal traffic 4 -> 2 block reads per
// entry and iteration.  Same mathematics, re-associated: 6e-16 against the reference-pinned oracle after 5
// iterations (tests/lusgs_fused_np.py restates exactly this order on the CPU).

// s[r] = sum_{c after r} U[r,c] x[c]   (RHSUx of SparseSolver.cpp:64-69, before the D^-1)
template <int B>
__global__ void k_ux_raw(int n, const int* Uptr, const int* Ucol, const int* Upos, const int* rmap, const double* val, const double* x, double* s) {
    const Grp<B> g(n);
    if (!g.on) return;
    const int r = g.row, i = g.i;
    double acc = 0.0;
    for (int k = Uptr[r]; k < Uptr[r + 1]; k++)
        acc += row_dot<B>(val + (size_t)Upos[k] * B * B + i * B, x + (size_t)rmap[Ucol[k]] * B);
    s[(size_t)r * B + i] = acc;
}

// one level of the fused forward sweep: v[r] = b[r] + sum_k LD[k] t[col[k]],  t[r] = s[r] - v[r]
template <int B>
__global__ void k_fwd_fused(int nrows, const int* rows, const int* ptr, const int* col, const int* rmap, const double* LD, const double* b,
                            const double* s, double* v, double* t) {
    const Grp<B> g(nrows);
    if (!g.on) return;
    const int r = rows[g.row], i = g.i;
    double acc = b[(size_t)(rmap ? rmap[r] : r) * B + i];  // rmap == nullptr: b is the solver's own beff (sweep numbering)
    for (int k = ptr[r]; k < ptr[r + 1]; k++) acc += row_dot<B>(LD + (size_t)k * B * B + i * B, t + (size_t)col[k] * B);
    v[(size_t)r * B + i] = acc;
    t[(size_t)r * B + i] = s[(size_t)r * B + i] - acc;
}

// one level of the backward sweep that also leaves s[r] = sum_k UD[k] w[col[k]] (= U x of the next iteration);
// w is updated exactly as k_sweep_level<B, false> does it
template <int B>
__global__ void k_bwd_fused(int nrows, const int* rows, const int* ptr, const int* col, const double* UD, double* w, double* s) {
    const Grp<B> g(nrows);
    if (!g.on) return;
    const int r = rows[g.row], i = g.i;
    double acc = w[(size_t)r * B + i], ss = 0.0;
    const int k0 = ptr[r], k1 = ptr[r + 1];
    for (int kk = 0; kk < k1 - k0; kk++) {
        const int k = k1 - 1 - kk;
        const double d = row_dot<B>(UD + (size_t)k * B * B + i * B, w + (size_t)col[k] * B);
        acc -= d;
        ss += d;
    }
    w[(size_t)r * B + i] = acc;
    s[(size_t)r * B + i] = ss;
}

// ---- lean iteration (mode 2) ----------------------------------------------------------------------------
// Mode 1 without the two whole-vector passes around the backward sweep: the reference's X1 = D^-1 rhs; rhs1 = D X1
// (SparseSolver.cpp:86-90) is the identity up to rounding (cond(D) eps), so the backward sweep starts from v
// itself, and x = D^-1 w is formed by the row's lane group the moment w[r] is final.  Per row and iteration this
// drops D and D^-1 of k_mid and a second read of w: ~2.0 -> ~1.5 kB.  Level 0 runs too (x = D^-1 v there).
template <int B>
__global__ void k_bwd_lean(int nrows, const int* rows, const int* ptr, const int* col, const int* rmap, const double* UD, const double* Dinv,
                           const double* v, double* w, double* s, double* x) {
    const Grp<B> g(nrows);
    const int r = g.on ? rows[g.row] : 0, i = g.i;
    double acc = 0.0;
    if (g.on) {
        double ss = 0.0;
        acc = v[(size_t)r * B + i];
        const int k0 = ptr[r], k1 = ptr[r + 1];
        for (int kk = 0; kk < k1 - k0; kk++) {
            const int k = k1 - 1 - kk;
            const double d = row_dot<B>(UD + (size_t)k * B * B + i * B, w + (size_t)col[k] * B);
            acc -= d;
            ss += d;
        }
        w[(size_t)r * B + i] = acc;
        s[(size_t)r * B + i] = ss;
    }
    if (B == 1) { if (g.on) x[rmap[r]] = Dinv[r] * acc; return; }
    const double xi = g.matvec_lanes(g.on ? Dinv + (size_t)r * B * B + i * B : nullptr, acc);
    if (g.on) x[(size_t)rmap[r] * B + i] = xi;
}

// X1 = D^-1 rhs; rhs1 = D X1   (SparseSolverNUM.cpp:184-187)
template <int B>
__global__ void k_mid(int n, const double* D, const double* Dinv, const double* rhs, double* rhs1) {
    const Grp<B> g(n);
    const int r = g.row, i = g.i;
    if (B == 1) { if (g.on) { const double x1 = (1. / D[r]) * rhs[r]; rhs1[r] = D[r] * x1; } return; }
    const double x1 = g.on ? row_dot<B>(Dinv + (size_t)r * B * B + i * B, rhs + (size_t)r * B) : 0.0;
    const double t = g.matvec_lanes(g.on ? D + (size_t)r * B * B + i * B : nullptr, x1);
    if (g.on) rhs1[(size_t)r * B + i] = t;
}

// xnew = D^-1 rhs1; residual (scalar only); x = xnew   (SparseSolverNUM.cpp:194-203)
template <int B>
__global__ void k_fin(int n, const int* rmap, const double* D, const double* Dinv, const double* rhs1, double* x, unsigned long long* res) {
    const Grp<B> g(n);
    const int r = g.row, i = g.i;
    double rr = 0.0;
    if (g.on) {
        const size_t rx = (size_t)rmap[r];
        if (B == 1) {
            const double xn = (1. / D[r]) * rhs1[r];
            const double q = fabs(x[rx] - xn) / x[rx];
            rr = (q > 0.0) ? q : 0.0;
            x[rx] = xn;
        } else {
            x[rx * B + i] = row_dot<B>(Dinv + (size_t)r * B * B + i * B, rhs1 + (size_t)r * B);
        }
    }
    if (B == 1) {
        for (int o = 16; o > 0; o >>= 1) rr = fmax(rr, __shfl_xor_sync(0xffffffffu, rr, o));
        if ((threadIdx.x & 31) == 0 && rr > 0.0) atomicMax(res, (unsigned long long)__double_as_longlong(rr));
    }
}

template <typename T>
int up(mstgpu_lusgs* h, T** d, const std::vector<T>& v) {
    LCK(cudaMalloc((void**)d, std::max<size_t>(1, v.size()) * sizeof(T)));
    if (!v.empty()) LCK(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// the sweeps on DEVICE arrays (val: caller's CSR order; x: start vector in, solution out)
template <int B>
int solve_core(mstgpu_lusgs* h, const double* val, const double* b, double* x, int max_iter, int early_exit,
               double* res_hist, int32_t* iters_done, bool setup = true) {
    const int n = h->n, T = 128;
    cudaStream_t s = h->stream;
    using G = Grp<B>;
    if (setup) {  // D, D^-1 and the scaled copies depend on the matrix only
        if (h->setup_mode == 0) {  // round-1 kernels: thread per row / per (entry, row), sweep order
            k_diag<B><<<(n + T - 1) / T, T, 0, s>>>(n, h->Dptr, h->Dpos, val, h->D, h->Dinv);
            if (h->nL) k_scale<B><<<(unsigned)(((size_t)h->nL * B + T - 1) / T), T, 0, s>>>((size_t)h->nL, h->Lcol, h->Lpos, val, h->D, h->Dinv, h->LD);
            if (h->nU) k_scale<B><<<(unsigned)(((size_t)h->nU * B + T - 1) / T), T, 0, s>>>((size_t)h->nU, h->Ucol, h->Upos, val, h->D, h->Dinv, h->UD);
            h->launches += 3;
        } else {  // LD / UD: lane group per block row, rows in the caller's order; D, D^-1: registers (2) or lane groups (1)
            if (h->setup_mode == 2) k_diag_reg<B><<<(n + 127) / 128, 128, 0, s>>>(n, h->Dptr, h->Dpos, val, h->D, h->Dinv);
            else k_diag_grp<B><<<G::grid(n, T), T, 0, s>>>(n, h->rinv, h->Dptr, h->Dpos, val, h->D, h->Dinv);
            k_scale_rows<B><<<G::grid(n, T), T, 0, s>>>(n, h->rinv, h->Lptr, h->Lcol, h->Lpos, h->Uptr, h->Ucol, h->Upos, val, h->D, h->Dinv, h->LD, h->UD);
            h->launches += 2;
        }
    }
    // the lean mode has no k_fin, which is where the scalar solver's residual (early exit) is formed
    const bool lean = h->mode == 2 && !(B == 1 && (res_hist || early_exit));
    const bool fused = h->mode == 1 || h->mode == 2;
    if (fused) {
        if (!h->s) LCK(cudaMalloc((void**)&h->s, (size_t)n * B * 8));
        if (setup) h->s_valid = false;
        if (h->x0_zero) {
            LCK(cudaMemsetAsync(h->s, 0, (size_t)n * B * 8, s));  // U 0 = 0
        } else if (!h->s_valid) {
            // the only read of the unscaled off-diagonal blocks in these modes: U x of the start vector
            k_ux_raw<B><<<G::grid(n, T), T, 0, s>>>(n, h->Uptr, h->Ucol, h->Upos, h->rmap, val, x, h->s);
            h->launches++;
        }
        h->x0_zero = false;
    }
    int it = 0;
    for (; it < max_iter; it++) {
        LCK(cudaMemsetAsync(h->res, 0, 8, s));
        const double* bb = b;
        const int* bmap = h->rmap;  // b is the caller's (row numbering); beff is the solver's own (sweep numbering)
        if (h->nG) {
            k_ghost_rhs<B><<<G::grid(n, T), T, 0, s>>>(n, h->Gptr, h->Gcol, h->Gpos, h->rmap, val, b, x, h->beff);
            h->launches++;
            bb = h->beff;
            bmap = nullptr;
        }
        if (fused) {
            // forward: every level, level 0 included (v = b there, but t = s - v is needed by the later levels);
            // h->rhs = v, h->ux = t
            for (size_t l = 0; l + 1 < h->fptr.size(); l++) {
                const int cnt = h->fptr[l + 1] - h->fptr[l];
                if (cnt > 0)
                    k_fwd_fused<B><<<G::grid(cnt, T), T, 0, s>>>(cnt, h->frows + h->fptr[l], h->Lptr, h->Lcol, bmap, h->LD, bb, h->s, h->rhs, h->ux);
            }
            if (lean) {
                for (size_t l = 0; l + 1 < h->bptr.size(); l++) {
                    const int cnt = h->bptr[l + 1] - h->bptr[l];
                    if (cnt > 0)
                        k_bwd_lean<B><<<G::grid(cnt, T), T, 0, s>>>(cnt, h->brows + h->bptr[l], h->Uptr, h->Ucol, h->rmap, h->UD, h->Dinv, h->rhs, h->rhs1, h->s, x);
                }
                h->launches += (int64_t)(h->fptr.size() > 1 ? h->fptr.size() - 1 : 0) + (int64_t)(h->bptr.size() > 1 ? h->bptr.size() - 1 : 0);
            } else {
            k_mid<B><<<G::grid(n, T), T, 0, s>>>(n, h->D, h->Dinv, h->rhs, h->rhs1);
            LCK(cudaMemsetAsync(h->s, 0, (size_t)n * B * 8, s));  // rows without upper entries (level 0): U x = 0
            for (size_t l = 1; l + 1 < h->bptr.size(); l++) {
                const int cnt = h->bptr[l + 1] - h->bptr[l];
                if (cnt > 0)
                    k_bwd_fused<B><<<G::grid(cnt, T), T, 0, s>>>(cnt, h->brows + h->bptr[l], h->Uptr, h->Ucol, h->UD, h->rhs1, h->s);
            }
            k_fin<B><<<G::grid(n, T), T, 0, s>>>(n, h->rmap, h->D, h->Dinv, h->rhs1, x, h->res);
            h->launches += 2 + (int64_t)(h->fptr.size() > 1 ? h->fptr.size() - 1 : 0) + (int64_t)(h->bptr.size() > 2 ? h->bptr.size() - 2 : 0);
            }
            h->s_valid = true;  // s = U x of the x just written (owned columns; ghost columns live in G)
        } else {
        k_ux<B><<<G::grid(n, T), T, 0, s>>>(n, h->Uptr, h->Ucol, h->Upos, h->rmap, val, h->D, h->Dinv, x, h->ux);
        k_rhs<B><<<G::grid(n, T), T, 0, s>>>(n, h->Lptr, h->Lcol, h->Lpos, bmap, val, bb, h->ux, h->rhs);
        for (size_t l = 1; l + 1 < h->fptr.size(); l++) {  // level 0 has no dependencies: nothing to subtract
            const int cnt = h->fptr[l + 1] - h->fptr[l];
            k_sweep_level<B, true><<<G::grid(cnt, T), T, 0, s>>>(cnt, h->frows + h->fptr[l], h->Lptr, h->Lcol, h->LD, h->rhs);
        }
        k_mid<B><<<G::grid(n, T), T, 0, s>>>(n, h->D, h->Dinv, h->rhs, h->rhs1);
        for (size_t l = 1; l + 1 < h->bptr.size(); l++) {
            const int cnt = h->bptr[l + 1] - h->bptr[l];
            k_sweep_level<B, false><<<G::grid(cnt, T), T, 0, s>>>(cnt, h->brows + h->bptr[l], h->Uptr, h->Ucol, h->UD, h->rhs1);
        }
        k_fin<B><<<G::grid(n, T), T, 0, s>>>(n, h->rmap, h->D, h->Dinv, h->rhs1, x, h->res);
        h->launches += 4 + (int64_t)(h->fptr.size() > 2 ? h->fptr.size() - 2 : 0) + (int64_t)(h->bptr.size() > 2 ? h->bptr.size() - 2 : 0);
        }
        if (B == 1 && (res_hist || early_exit)) {
            unsigned long long bits = 0;
            LCK(cudaMemcpyAsync(&bits, h->res, 8, cudaMemcpyDeviceToHost, s));
            LCK(cudaStreamSynchronize(s));
            double r;
            std::memcpy(&r, &bits, 8);
            if (res_hist) res_hist[it] = r;
            if (early_exit && r > 1e-10 * 1e-10 && r < 1e-7) { it++; break; }  // SparseSolverNUM.cpp:205
        }
    }
    LCK(cudaGetLastError());
    if (iters_done) *iters_done = it;
    return MSTGPU_OK;
}

// host arrays in, host x out (the reference's calling convention: setELE / setD / setRHSb ... getPNewX)
template <int B>
int solve_impl(mstgpu_lusgs* h, const double* val, const double* b, double* x, int max_iter, int early_exit,
               double* res_hist, int32_t* iters_done) {
    const int n = h->n, BB = B * B;
    cudaStream_t s = h->stream;
    if (!h->val) {
        LCK(cudaMalloc((void**)&h->val, std::max<size_t>(1, h->nnz) * BB * 8));
        LCK(cudaMalloc((void**)&h->b, (size_t)n * B * 8));
        LCK(cudaMalloc((void**)&h->x, (size_t)h->ncols * B * 8));
    }
    LCK(cudaMemcpyAsync(h->val, val, (size_t)h->nnz * BB * 8, cudaMemcpyHostToDevice, s));
    LCK(cudaMemcpyAsync(h->b, b, (size_t)n * B * 8, cudaMemcpyHostToDevice, s));
    LCK(cudaMemcpyAsync(h->x, x, (size_t)h->ncols * B * 8, cudaMemcpyHostToDevice, s));  // ghost rows: the caller's values
    int rc = solve_core<B>(h, h->val, h->b, h->x, max_iter, early_exit, res_hist, iters_done);
    if (rc) return rc;
    LCK(cudaMemcpyAsync(x, h->x, (size_t)n * B * 8, cudaMemcpyDeviceToHost, s));
    LCK(cudaStreamSynchronize(s));
    return MSTGPU_OK;
}

}  // namespace

// in-library entry for the implicit step of mstgpu.cu: the sweeps on the caller's stream, no sync
namespace mst {
// the next solve starts from x = 0 (the implicit step's dQ): U x = 0 needs no pass over the U blocks
void lusgs_hint_zero_start(mstgpu_lusgs* h) { if (h) h->x0_zero = true; }
int lusgs_solve_async(mstgpu_lusgs* h, cudaStream_t st, const double* val, const double* b, double* x, int iters, bool setup) {
    cudaStream_t own = h->stream;
    h->stream = st;
    int rc;
    switch (h->B) {
        case 1: rc = solve_core<1>(h, val, b, x, iters, 0, nullptr, nullptr, setup); break;
        case 4: rc = solve_core<4>(h, val, b, x, iters, 0, nullptr, nullptr, setup); break;
        default: rc = solve_core<5>(h, val, b, x, iters, 0, nullptr, nullptr, setup); break;
    }
    h->stream = own;
    return rc;
}
}  // namespace mst

extern "C" {

const char* mstgpu_lusgs_last_error(void) { return g_lusgs_error.c_str(); }

// greedy first-fit colouring of the symmetrised pattern, rows visited in storage order; O(nnz), CSR only
static int color_rows(int32_t n, const int32_t* rowptr, const int32_t* col, std::vector<int>& color, int& nc, int32_t ncols = -1) {
    if (ncols < n) ncols = n;  // columns in [n, ncols) are ghost columns: no row to colour against
    // transpose pattern (for unsymmetric input): tptr / tcol
    std::vector<int64_t> tptr((size_t)n + 1, 0);
    for (int r = 0; r < n; r++)
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
            const int c = col[k];
            if (c < 0 || c >= ncols) { g_lusgs_error = "column out of range"; return MSTGPU_ERR_ARG; }
            if (c != r && c < n) tptr[(size_t)c + 1]++;
        }
    for (int r = 0; r < n; r++) tptr[r + 1] += tptr[r];
    std::vector<int> tcol((size_t)tptr[n]);
    {
        std::vector<int64_t> pos(tptr.begin(), tptr.end() - 1);
        for (int r = 0; r < n; r++)
            for (int k = rowptr[r]; k < rowptr[r + 1]; k++)
                if (col[k] != r && col[k] < n) tcol[(size_t)pos[col[k]]++] = r;
    }
    color.assign(n, -1);
    nc = 0;
    unsigned long long used;  // colours 0..63 as a bit mask (a mesh graph needs a handful)
    for (int r = 0; r < n; r++) {
        used = 0;
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++)
            if (col[k] != r && col[k] < n && color[col[k]] >= 0 && color[col[k]] < 64) used |= 1ULL << color[col[k]];
        for (int64_t k = tptr[r]; k < tptr[r + 1]; k++)
            if (color[tcol[(size_t)k]] >= 0 && color[tcol[(size_t)k]] < 64) used |= 1ULL << color[tcol[(size_t)k]];
        int k = 0;
        while (k < 63 && ((used >> k) & 1)) k++;
        color[r] = k;
        nc = std::max(nc, k + 1);
    }
    return MSTGPU_OK;
}

int mstgpu_lusgs_color_order(int32_t n, const int32_t* rowptr, const int32_t* col, int32_t* perm_new2old,
                             int32_t* ncolors) {
    return mstgpu_lusgs_color_order_partitioned(n, n, rowptr, col, perm_new2old, ncolors);
}

int mstgpu_lusgs_color_order_partitioned(int32_t n, int32_t ncols, const int32_t* rowptr, const int32_t* col,
                                         int32_t* perm_new2old, int32_t* ncolors) {
    if (n <= 0 || !rowptr || !col || !perm_new2old) { g_lusgs_error = "bad argument"; return MSTGPU_ERR_ARG; }
    std::vector<int> color;
    int nc = 0;
    int rc = color_rows(n, rowptr, col, color, nc, ncols);
    if (rc) return rc;
    // counting sort by colour (stable: storage order within a colour)
    std::vector<int64_t> start((size_t)nc + 1, 0);
    for (int r = 0; r < n; r++) start[(size_t)color[r] + 1]++;
    for (int c = 0; c < nc; c++) start[c + 1] += start[c];
    for (int r = 0; r < n; r++) perm_new2old[start[color[r]]++] = r;
    if (ncolors) *ncolors = nc;
    return MSTGPU_OK;
}

int mstgpu_lusgs_create(mstgpu_lusgs** out, int32_t n, int32_t block, const int32_t* rowptr, const int32_t* col,
                        int32_t device) {
    return mstgpu_lusgs_create_ordered(out, n, block, rowptr, col, nullptr, device);
}

int mstgpu_lusgs_create_ordered(mstgpu_lusgs** out, int32_t n, int32_t block, const int32_t* rowptr, const int32_t* col,
                                const int32_t* sweep_new2old, int32_t device) {
    return mstgpu_lusgs_create_partitioned(out, n, n, block, rowptr, col, sweep_new2old, device);
}

int mstgpu_lusgs_create_partitioned(mstgpu_lusgs** out, int32_t n, int32_t ncols, int32_t block, const int32_t* rowptr,
                                    const int32_t* col, const int32_t* sweep_new2old, int32_t device) {
    mstgpu_lusgs* h = nullptr;
    if (ncols < n) { g_lusgs_error = "ncols < n"; return MSTGPU_ERR_ARG; }
    if (!out || n <= 0 || !rowptr || !col) { g_lusgs_error = "bad argument"; return MSTGPU_ERR_ARG; }
    *out = nullptr;
    if (block != 1 && block != 4 && block != 5) { g_lusgs_error = "block size must be 1, 4 or 5"; return MSTGPU_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_lusgs_error = "no CUDA device"; return MSTGPU_ERR_CUDA; }
    // rank[r] = position of row r in the sweep; the sweeps are those of the reference on P A P^T, data unmoved
    std::vector<int> rank;
    if (sweep_new2old) {
        rank.assign(n, -1);
        for (int i = 0; i < n; i++) {
            const int r = sweep_new2old[i];
            if (r < 0 || r >= n || rank[r] >= 0) { g_lusgs_error = "sweep order is not a permutation"; return MSTGPU_ERR_ARG; }
            rank[r] = i;
        }
    }
    auto rk = [&](int r) { return sweep_new2old ? rank[r] : r; };
    // index lists in SWEEP numbering: row p = sweep position, columns = sweep positions of the coupled rows
    // (ghost columns >= n keep their index), entries of a row by ascending sweep position as the reference
    // subtracts them
    std::vector<int> Lptr(n + 1, 0), Uptr(n + 1, 0), Dptr(n + 1, 0), Gptr(n + 1, 0), Lcol, Lpos, Ucol, Upos, Dpos, Gcol, Gpos, rmap(n);
    std::vector<std::pair<int, int>> lo, hi;  // (sweep position of the column, index into the row)
    for (int r = 0; r < n; r++)
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
            const int c = col[k];
            if (c < 0 || c >= ncols) { g_lusgs_error = "column out of range"; return MSTGPU_ERR_ARG; }
            if (k > rowptr[r] && col[k - 1] > c) { g_lusgs_error = "columns must be ascending within a row"; return MSTGPU_ERR_ARG; }
        }
    for (int p = 0; p < n; p++) {
        const int r = sweep_new2old ? sweep_new2old[p] : p;
        rmap[p] = r;
        lo.clear(); hi.clear();
        for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
            const int c = col[k];
            if (c >= n) { Gcol.push_back(c); Gpos.push_back(k); }
            else if (c == r) Dpos.push_back(k);
            else if (rk(c) < p) lo.push_back({rk(c), k});
            else hi.push_back({rk(c), k});
        }
        if (sweep_new2old) { std::sort(lo.begin(), lo.end()); std::sort(hi.begin(), hi.end()); }
        for (auto& e : lo) { Lcol.push_back(e.first); Lpos.push_back(e.second); }
        for (auto& e : hi) { Ucol.push_back(e.first); Upos.push_back(e.second); }
        Lptr[p + 1] = (int)Lcol.size(); Uptr[p + 1] = (int)Ucol.size(); Dptr[p + 1] = (int)Dpos.size();
        Gptr[p + 1] = (int)Gcol.size();
        if (Dptr[p + 1] == Dptr[p]) { g_lusgs_error = "row without a diagonal entry"; return MSTGPU_ERR_ARG; }
    }
    // dependency levels, rows visited in sweep order (= ascending internal index)
    std::vector<int> lf(n, 0), lb(n, 0);
    int nlf = 0, nlb = 0;
    for (int p = 0; p < n; p++) {
        int l = 0;
        for (int k = Lptr[p]; k < Lptr[p + 1]; k++) l = std::max(l, lf[Lcol[k]] + 1);
        lf[p] = l; nlf = std::max(nlf, l + 1);
    }
    for (int p = n - 1; p >= 0; p--) {
        int l = 0;
        for (int k = Uptr[p]; k < Uptr[p + 1]; k++) l = std::max(l, lb[Ucol[k]] + 1);
        lb[p] = l; nlb = std::max(nlb, l + 1);
    }
    auto bucket = [&](const std::vector<int>& lev, int nl, std::vector<int>& ptr, std::vector<int>& rows) {
        ptr.assign(nl + 1, 0);
        for (int r = 0; r < n; r++) ptr[lev[r] + 1]++;
        for (int l = 0; l < nl; l++) ptr[l + 1] += ptr[l];
        rows.resize(n);
        std::vector<int> pos(ptr.begin(), ptr.end() - 1);
        for (int r = 0; r < n; r++) rows[pos[lev[r]]++] = r;  // storage order within a level
    };
    h = new mstgpu_lusgs;
    h->n = n; h->B = block; h->nnz = rowptr[n]; h->nL = (int)Lcol.size(); h->nU = (int)Ucol.size();
    h->nG = (int)Gcol.size(); h->ncols = ncols;
    std::vector<int> frows, brows;
    bucket(lf, nlf, h->fptr, frows);
    bucket(lb, nlb, h->bptr, brows);
    int rc = [&]() -> int {
        if (device >= 0) LCK(cudaSetDevice(device));
        LCK(cudaGetDevice(&h->device));
        LCK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        LCK(cudaEventCreate(&h->ev0));
        LCK(cudaEventCreate(&h->ev1));
        int r;
        if ((r = up(h, &h->Lptr, Lptr)) || (r = up(h, &h->Lcol, Lcol)) || (r = up(h, &h->Lpos, Lpos))) return r;
        if ((r = up(h, &h->Uptr, Uptr)) || (r = up(h, &h->Ucol, Ucol)) || (r = up(h, &h->Upos, Upos))) return r;
        if ((r = up(h, &h->Dptr, Dptr)) || (r = up(h, &h->Dpos, Dpos))) return r;
        if (h->nG) {
            if ((r = up(h, &h->Gptr, Gptr)) || (r = up(h, &h->Gcol, Gcol)) || (r = up(h, &h->Gpos, Gpos))) return r;
            LCK(cudaMalloc((void**)&h->beff, (size_t)n * block * 8));
        }
        if ((r = up(h, &h->frows, frows)) || (r = up(h, &h->brows, brows)) || (r = up(h, &h->rmap, rmap))) return r;
        {
            std::vector<int> rinv(n);
            for (int q = 0; q < n; q++) rinv[rmap[q]] = q;
            if ((r = up(h, &h->rinv, rinv))) return r;
        }
        const size_t BB = (size_t)block * block;
        // val / b / x buffers of the host-array entry point are allocated at its first use
        LCK(cudaMalloc((void**)&h->D, n * BB * 8));
        LCK(cudaMalloc((void**)&h->Dinv, n * BB * 8));
        LCK(cudaMalloc((void**)&h->LD, std::max<size_t>(1, h->nL) * BB * 8));
        LCK(cudaMalloc((void**)&h->UD, std::max<size_t>(1, h->nU) * BB * 8));
        for (double** p : {&h->rhs, &h->rhs1, &h->ux}) LCK(cudaMalloc((void**)p, (size_t)n * block * 8));
        LCK(cudaMalloc((void**)&h->res, 8));
        return 0;
    }();
    if (rc) { mstgpu_lusgs_destroy(h); return rc; }
    if (const char* v = getenv("MSTGPU_LUSGS_MODE")) { const int m = atoi(v); h->mode = (m >= 0 && m <= 2) ? m : 0; }
    if (const char* v = getenv("MSTGPU_LUSGS_SETUP")) { const int m = atoi(v); h->setup_mode = (m >= 0 && m <= 2) ? m : 2; }
    *out = h;
    return MSTGPU_OK;
}

void mstgpu_lusgs_destroy(mstgpu_lusgs* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : {(void*)h->Lptr, (void*)h->Lcol, (void*)h->Lpos, (void*)h->Uptr, (void*)h->Ucol, (void*)h->Upos,
                    (void*)h->Dptr, (void*)h->Dpos, (void*)h->frows, (void*)h->brows, (void*)h->val, (void*)h->D,
                    (void*)h->Dinv, (void*)h->LD, (void*)h->UD, (void*)h->b, (void*)h->x, (void*)h->rhs, (void*)h->rhs1,
                    (void*)h->ux, (void*)h->s, (void*)h->res, (void*)h->Gptr, (void*)h->Gcol, (void*)h->Gpos, (void*)h->beff, (void*)h->rmap, (void*)h->rinv})
        if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int mstgpu_lusgs_levels(mstgpu_lusgs* h, int32_t* fwd, int32_t* bwd) {
    if (!h) return MSTGPU_ERR_ARG;
    if (fwd) *fwd = (int32_t)h->fptr.size() - 1;
    if (bwd) *bwd = (int32_t)h->bptr.size() - 1;
    return MSTGPU_OK;
}

int mstgpu_lusgs_solve(mstgpu_lusgs* h, const double* val, const double* b, double* x, int32_t max_iter,
                       int32_t early_exit, double* res_hist, int32_t* iters_done) {
    if (!h || !val || !b || !x || max_iter < 0) { g_lusgs_error = "bad argument"; return MSTGPU_ERR_ARG; }
    LCK(cudaSetDevice(h->device));
    switch (h->B) {
        case 1: return solve_impl<1>(h, val, b, x, max_iter, early_exit, res_hist, iters_done);
        case 4: return solve_impl<4>(h, val, b, x, max_iter, early_exit, res_hist, iters_done);
        default: return solve_impl<5>(h, val, b, x, max_iter, early_exit, res_hist, iters_done);
    }
}

int mstgpu_lusgs_solve_device(mstgpu_lusgs* h, const double* d_val, const double* d_b, double* d_x, int32_t max_iter,
                              float* ms) {
    if (!h || !d_val || !d_b || !d_x || max_iter < 0) { g_lusgs_error = "bad argument"; return MSTGPU_ERR_ARG; }
    LCK(cudaSetDevice(h->device));
    LCK(cudaStreamSynchronize(h->stream));
    if (ms) LCK(cudaEventRecord(h->ev0, h->stream));
    int rc;
    switch (h->B) {
        case 1: rc = solve_core<1>(h, d_val, d_b, d_x, max_iter, 0, nullptr, nullptr); break;
        case 4: rc = solve_core<4>(h, d_val, d_b, d_x, max_iter, 0, nullptr, nullptr); break;
        default: rc = solve_core<5>(h, d_val, d_b, d_x, max_iter, 0, nullptr, nullptr); break;
    }
    if (rc) return rc;
    if (ms) {
        LCK(cudaEventRecord(h->ev1, h->stream));
        LCK(cudaEventSynchronize(h->ev1));
        LCK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    } else {
        LCK(cudaStreamSynchronize(h->stream));
    }
    return MSTGPU_OK;
}

int64_t mstgpu_lusgs_launch_count(mstgpu_lusgs* h) { return h ? h->launches : -1; }

int64_t mstgpu_lusgs_device_bytes(mstgpu_lusgs* h) {
    if (!h) return -1;
    const int64_t BB = (int64_t)h->B * h->B;
    return ((int64_t)h->n * 2 + h->nL + h->nU) * BB * 8 + (int64_t)h->n * h->B * 8 * (3 + (h->s ? 1 : 0)) +
           ((int64_t)h->n * 5 + 3 + 2LL * (h->nL + h->nU)) * 4;
}

int mstgpu_lusgs_set_mode(mstgpu_lusgs* h, int32_t mode) {
    if (!h || mode < 0 || mode > 2) { g_lusgs_error = "mode must be 0 (reference passes), 1 (fused) or 2 (lean)"; return MSTGPU_ERR_ARG; }
    h->mode = mode;
    h->s_valid = false;
    return MSTGPU_OK;
}

}  // extern "C"
