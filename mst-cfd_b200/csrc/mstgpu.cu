// mstgpu.cu -- C ABI (include/mstgpu.h) and the sm_100a kernels of the
// rhoSolver hot path.  Reference path being replaced (R = /root/reference/MST-CFD):
//   K1+K2  face interpolation + Green-Gauss gradient   R/rhoSolver/RhoSolver.cpp:430-452
//   K3     reconstruction + Roe/AUSM+ face flux        RhoSolver.cpp:90-369, SolverRoe.cpp, SolverAusm.cpp
//   K4+K5  cell gather + explicit Euler + residual     RhoSolver.cpp:45-68, R/time/Time.cpp:69-76
//   K6     new -> old                                  RhoSolver.cpp:513-517 (pointer swap here)
//
// No CPU fallback exists in this file: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/mstgpu.h"
#include "physics.cuh"
#include "partition.h"
#include "plan.h"
#include "step_tiles.cuh"
#include "tiles.h"

using namespace mst;

namespace mst {
// csrc/lusgs.cu: the sweeps on the caller's stream, device arrays, no synchronisation
int lusgs_solve_async(mstgpu_lusgs* h, cudaStream_t st, const double* val, const double* b, double* x, int iters, bool setup);
void lusgs_hint_zero_start(mstgpu_lusgs* h);
}

#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(e_));        \
            return MSTGPU_ERR_CUDA;                                                    \
        }                                                                              \
    } while (0)

namespace {
thread_local std::string g_create_error;

// NCCL is bound at run time, on first use, not at link time: a host process
// that also runs PyTorch already carries a libnccl.so.2 of its own (newer than
// the system one), and two different NCCLs under one soname cannot coexist.
// dlopen by soname returns whichever copy the process has loaded already.
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define MST_SYM(field, name)                                                   \
    field = reinterpret_cast<decltype(field)>(dlsym(h, name));                 \
    if (!field) { err = std::string("libnccl lacks ") + name; h = nullptr; return false; }
        MST_SYM(GetUniqueId, "ncclGetUniqueId")
        MST_SYM(CommInitRank, "ncclCommInitRank")
        MST_SYM(CommDestroy, "ncclCommDestroy")
        MST_SYM(GroupStart, "ncclGroupStart")
        MST_SYM(GroupEnd, "ncclGroupEnd")
        MST_SYM(Send, "ncclSend")
        MST_SYM(Recv, "ncclRecv")
        MST_SYM(AllReduce, "ncclAllReduce")
        MST_SYM(GetErrorString, "ncclGetErrorString")
#undef MST_SYM
        return true;
    }
};
NcclApi g_nccl;
}

#define MSTGPU_MAX_NB 32
struct PeerTable {
    double* q[2][MSTGPU_MAX_NB];
    unsigned long long* flag[MSTGPU_MAX_NB];
    int nnb;
};

struct KernelStat {
    double ms = 0.0;
    int64_t launches = 0;
};

struct mstgpu_ctx {
    Plan plan;  // host copy of permutations (tables are freed after upload)
    mstgpu_config cfg;
    DevCfg dcfg;
    int device = 0;
    int D = 0, U = 0, nc = 0, nf = 0, nslot = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // device tables
    double *Q[2] = {nullptr, nullptr}, *G = nullptr, *Gp = nullptr, *Phi = nullptr, *stage = nullptr;
    double *Sd = nullptr, *dx0 = nullptr, *dx1 = nullptr, *eta = nullptr, *vol = nullptr;
    int32_t *fc0 = nullptr, *fc1 = nullptr, *cf = nullptr, *cell_new2old = nullptr, *face_new2old = nullptr;
    uint32_t* meta = nullptr;
    // fused tile kernel
    TileArrays ta{};
    std::vector<void*> tile_allocs;
    int ntiles = 0, tile_T = 0, tile_NT = 0;
    int tile_var = 0;   // MSTGPU_TILE_VAR: experimental variant of the default fused instantiation (0 = default)
    std::map<const void*, size_t> smem_configured;  // dynamic shared memory opted in per kernel instantiation, on this context's device
    bool tile_staged = false;  // packet stream staged through shared memory (step_tiles.cuh, VAR & 128)
    int tile_ext = 0;          // tile_layout() ext argument of this context's tiles
    int ring_stride = 0;       // ring ids per tile in the device array (fixed stride)
    int sm_count = 148;
    size_t tile_smem = 0;
    struct TileClass { int first, count; size_t smem; bool halo; };  // halo: a ring of the tile holds ghost cells
    std::vector<TileClass> tile_classes;  // tiles grouped by shared-memory need (CTAs per SM)
    bool use_tiles = false;
    bool split_ready = false;   // the per-face / per-cell tables of the split kernels are on the device (uploaded on first use
                                // when the fused kernel is the step path: it never reads them)
    size_t stage_cap = 0;       // doubles in `stage`
    bool probes_valid = false;  // G / Phi hold the stages of the last step
    // multi-GPU (partitioned context)
    int n_owned = 0;  // cells advanced by this context (== nc when not partitioned)
    bool partitioned = false;
    struct HaloNb { int rank, send_off, send_count, recv_first, recv_count; };
    std::vector<HaloNb> halo;
    int32_t* send_idx = nullptr;  // device-order ids of the owned cells to send, all neighbours back to back
    double* sendbuf = nullptr;
    int send_total = 0;
    cudaStream_t stream2 = nullptr;  // halo exchange, overlapped with the interior tiles
    cudaEvent_t ev_halo = nullptr, ev_done = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // peer-memory halo (mstgpu_peer_connect): boundary rows are STORED into the neighbours' ghost blocks over
    // NVLink by this rank's own kernel, followed by a release flag; no pack buffer, no NCCL kernel on the data path
    bool peer_ok = false;
    PeerTable peer{};                        // neighbours' Q[0] / Q[1] and the address of MY slot in their flag arrays
    unsigned long long* peer_flags = nullptr;  // [MSTGPU_MAX_NB] epochs written by the neighbours (slot = index in halo[])
    int32_t* push_dst = nullptr;             // [send_total] row in the receiver's local numbering
    uint8_t* push_slot = nullptr;            // [send_total] neighbour index
    unsigned int* push_ticket = nullptr;     // last-CTA-done counter of k_peer_push
    unsigned long long* peer_epochs = nullptr;  // device: [0] exchanges pushed, [1] exchanges received (every rank runs the same sequence)
    std::vector<void*> peer_opened;          // cudaIpcOpenMemHandle results (closed in destroy)
    // multi-step CUDA graph: two ping-pong steps, the tile classes of a step as parallel
    // branches; one executable per starting buffer, rebuilt when dt changes
    cudaGraphExec_t step_graph[2] = {nullptr, nullptr};
    double step_graph_dt = 0.0;
    cudaStream_t fork_stream = nullptr;  // non-null while a step is captured: odd classes go here
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int64_t graph_launches = 0;
    int graph_rc = 0;  // result of the last graph build
    // implicit step (extension): block system + the reference's LU-SGS sweeps (csrc/lusgs.cu)
    mstgpu_lusgs* imp_solver = nullptr;
    int32_t *imp_dpos = nullptr, *imp_pos = nullptr;
    double *imp_val = nullptr, *imp_b = nullptr, *imp_x = nullptr;
    std::vector<int32_t> imp_sweep;  // sweep order, device cell ids (empty = storage order)
    // extension tables / CFL stepping (mstgpu_config.gradient / limiter, mstgpu_step_cfl)
    double *lsq = nullptr, *eps2 = nullptr;
    unsigned long long* dtmin = nullptr;  // bit pattern of the smallest cell time step
    double* dt_dev = nullptr;             // [0] dt of the running step, [1] time advanced
    unsigned long long* resid = nullptr;  // [U] bit patterns of non-negative doubles
    int* nanflag = nullptr;
    // output path (csrc/output.cuh): node -> faces CSR, per-face cells (device order) and eta, node weights
    int32_t *out_nf_ptr = nullptr, *out_nf_idx = nullptr, *out_c0 = nullptr, *out_c1 = nullptr;
    double *out_eta = nullptr, *out_w = nullptr, *out_fields = nullptr;
    int out_nn = 0;
    // partitioned output: the nodes this rank computes, the fresh rows it needs from other ranks and owes them
    std::vector<int32_t> out_node_ids;        // global ids of the nodes of out_fields, ascending (empty = all nodes, in order)
    double* out_Q = nullptr;                  // [nc + out_nextra][U]: copy of the local state + rows received for the output
    int out_nextra = 0;
    struct OutNb { int rank, send_off, send_count, recv_off, recv_count; };
    std::vector<OutNb> out_nbrs;
    int32_t* out_send_idx = nullptr;          // device-order local ids of the owned rows to send, all ranks back to back
    double* out_sendbuf = nullptr;
    int out_send_total = 0;
    // streamed step (mstgpu_step_host): host rows in, host rows out, pipelined over chunks of host rows
    std::vector<int32_t> tile_ready_row;          // per tile (desc[] order): largest host row among its owned + owned-range ring cells
    std::vector<int32_t> tile_cb_h, tile_nown_h;  // host copy of the descriptors' owned ranges
    struct HostStream {
        int nchunks = 0;
        int64_t chunk_rows = 0;
        int32_t *d_order = nullptr, *d_old2new = nullptr;
        std::vector<std::vector<int>> off;         // [class][group 0 .. nchunks + 1]: first entry of the group in d_order
        std::vector<std::vector<int>> out_chunks;  // [group 0 .. nchunks]: host chunks complete once that group has run
        cudaStream_t s_in = nullptr, s_out = nullptr;
        std::vector<cudaEvent_t> ev_in, ev_out;
        cudaEvent_t ev_start = nullptr;
    } hs;
    int sub_group = -1;  // >= 0 while mstgpu_step_host launches the tiles of one group
    int cur = 0;          // Q[cur] = current ("old") state
    bool has_state = false, stepped = false;
    int64_t launches = 0, dev_bytes = 0;
    bool ktiming = false;
    std::map<std::string, KernelStat> kstat;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
    std::vector<cudaEvent_t> evpool;
    std::string err;
};

static void set_error(mstgpu_ctx* ctx, const std::string& s) {
    if (ctx) ctx->err = s;
    else g_create_error = s;
}

// ============================================================================
// kernels (V1: one thread per cell / face, global-memory gathers)
// ============================================================================
namespace {

template <int D>
__global__ void __launch_bounds__(256) k_gradient(int nc, int nslot, const double* __restrict__ Q,
                                                  const int32_t* __restrict__ cf,
                                                  const int32_t* __restrict__ fc0,
                                                  const int32_t* __restrict__ fc1,
                                                  const double* __restrict__ eta,
                                                  const double* __restrict__ Sd,
                                                  const double* __restrict__ vol,
                                                  double* __restrict__ G, double* __restrict__ Gp, double cv) {
    constexpr int U = D + 2;
    constexpr int P = D + 1;  // primitives differentiated for the viscous term: u_0..u_{D-1}, T
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    double tp[P][D];
#pragma unroll
    for (int k = 0; k < P; k++)
#pragma unroll
        for (int d = 0; d < D; d++) tp[k][d] = 0.0;
    double qc[U];
#pragma unroll
    for (int k = 0; k < U; k++) qc[k] = Q[(size_t)c * U + k];
    double t[U][D];
#pragma unroll
    for (int k = 0; k < U; k++)
#pragma unroll
        for (int d = 0; d < D; d++) t[k][d] = 0.0;
    for (int j = 0; j < nslot; j++) {
        const int v = cf[(size_t)j * nc + c];
        if (v < 0) continue;
        const int f = v >> 1;
        const int side = v & 1;
        const double e = eta[f];
        const int nb = side ? fc0[f] : fc1[f];
        double qf[U];
        if (nb >= 0) {
            // Qf = eta*Q[c0] + (1-eta)*Q[c1]   (RhoSolver.cpp:435)
            const double e0 = side ? (1.0 - e) : e;  // weight of this cell
            const double e1 = side ? e : (1.0 - e);  // weight of the neighbour
#pragma unroll
            for (int k = 0; k < U; k++) qf[k] = e0 * qc[k] + e1 * Q[(size_t)nb * U + k];
        } else {
#pragma unroll
            for (int k = 0; k < U; k++) qf[k] = qc[k];  // RhoSolver.cpp:439
        }
        const double sg = side ? -1.0 : 1.0;  // outward from this cell (MshBlock.cpp:307-318)
#pragma unroll
        for (int d = 0; d < D; d++) {
            const double s = sg * Sd[(size_t)f * D + d];
#pragma unroll
            for (int k = 0; k < U; k++) t[k][d] += qf[k] * s;
        }
        if (Gp) {
            // Green-Gauss gradient of the face primitives (laminar viscous extension)
            double pr[P], m2 = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) { pr[i] = qf[i + 1] / qf[0]; m2 += qf[i + 1] * qf[i + 1]; }
            pr[D] = (qf[U - 1] - 0.5 * m2 / qf[0]) / qf[0] / cv;  // FUNCTION.cpp:8-11
#pragma unroll
            for (int d = 0; d < D; d++) {
                const double s = sg * Sd[(size_t)f * D + d];
#pragma unroll
                for (int k = 0; k < P; k++) tp[k][d] += pr[k] * s;
            }
        }
    }
    const double v = vol[c];
    if (Gp) {
#pragma unroll
        for (int k = 0; k < P; k++)
#pragma unroll
            for (int d = 0; d < D; d++) Gp[((size_t)c * P + k) * D + d] = tp[k][d] / v;
    }
#pragma unroll
    for (int k = 0; k < U; k++)
#pragma unroll
        for (int d = 0; d < D; d++) G[((size_t)c * U + k) * D + d] = t[k][d] / v;
}

template <int D, int ORDER>
__global__ void __launch_bounds__(128) k_flux(int nf, DevCfg cfg, const double* __restrict__ Q,
                                              const double* __restrict__ G,
                                              const int32_t* __restrict__ fc0,
                                              const int32_t* __restrict__ fc1,
                                              const uint32_t* __restrict__ meta,
                                              const double* __restrict__ Sd,
                                              const double* __restrict__ dx0,
                                              const double* __restrict__ dx1,
                                              double* __restrict__ Phi, const double* __restrict__ Gp,
                                              const double* __restrict__ eta) {
    constexpr int U = D + 2;
    constexpr int P = D + 1;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int a = fc0[f], b = fc1[f];
    const uint32_t mt = meta[f];
    const int type = mt & 0xff;
    const uint32_t flags = mt >> 8;
    double S[D];
#pragma unroll
    for (int d = 0; d < D; d++) S[d] = Sd[(size_t)f * D + d];
    double qa[U], ra[U];
#pragma unroll
    for (int k = 0; k < U; k++) qa[k] = Q[(size_t)a * U + k];
    if (ORDER == 2) {
        double dx[D];
#pragma unroll
        for (int d = 0; d < D; d++) dx[d] = dx0[(size_t)f * D + d];
#pragma unroll
        for (int k = 0; k < U; k++) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < D; d++) s += G[((size_t)a * U + k) * D + d] * dx[d];
            ra[k] = qa[k] + s;
        }
    } else {
#pragma unroll
        for (int k = 0; k < U; k++) ra[k] = qa[k];
    }
    double A[U], B[U], phi[U];
    bool live = true;
    if (b >= 0) {
#pragma unroll
        for (int k = 0; k < U; k++) A[k] = ra[k];
        if (ORDER == 2) {
            double dx[D];
#pragma unroll
            for (int d = 0; d < D; d++) dx[d] = dx1[(size_t)f * D + d];
#pragma unroll
            for (int k = 0; k < U; k++) {
                double s = 0.0;
#pragma unroll
                for (int d = 0; d < D; d++) s += G[((size_t)b * U + k) * D + d] * dx[d];
                B[k] = Q[(size_t)b * U + k] + s;
            }
        } else {
#pragma unroll
            for (int k = 0; k < U; k++) B[k] = Q[(size_t)b * U + k];
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
    if (Gp) {
        // laminar viscous flux, corrected formulation (the reference's updateViscid,
        // RhoSolver.cpp:371-429, cannot run): tau = mu (grad u + grad u^T) + lambda div(u) I,
        // energy flux u.tau + k grad T, face values by the eta interpolation of the gradient pass
        const double e = eta[f];
        double qf[U], gf[P][D];
        if (b >= 0) {
#pragma unroll
            for (int k = 0; k < U; k++) qf[k] = e * qa[k] + (1.0 - e) * Q[(size_t)b * U + k];
#pragma unroll
            for (int k = 0; k < P; k++)
#pragma unroll
                for (int d = 0; d < D; d++)
                    gf[k][d] = e * Gp[((size_t)a * P + k) * D + d] + (1.0 - e) * Gp[((size_t)b * P + k) * D + d];
        } else {
#pragma unroll
            for (int k = 0; k < U; k++) qf[k] = qa[k];
#pragma unroll
            for (int k = 0; k < P; k++)
#pragma unroll
                for (int d = 0; d < D; d++) gf[k][d] = Gp[((size_t)a * P + k) * D + d];
        }
        double uf[D], div = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) { uf[i] = qf[i + 1] / qf[0]; div += gf[i][i]; }
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
#pragma unroll
        for (int i = 0; i < D; i++) phi[i + 1] -= fv[i];
        phi[U - 1] -= en;
    }
#pragma unroll
    for (int k = 0; k < U; k++) Phi[(size_t)f * U + k] = phi[k];
}

// ---- extension kernels (absent from the reference; include/mstgpu.h, mstgpu_config) --------------

// least-squares gradient with the fixed weights of plan.h: G_c = sum_j lsq[j][:][c] (Q_nb(j) - Q_c)
template <int D>
__global__ void __launch_bounds__(256) k_gradient_lsq(int nc, int nslot, const double* __restrict__ Q,
                                                      const int32_t* __restrict__ cf,
                                                      const int32_t* __restrict__ fc0,
                                                      const int32_t* __restrict__ fc1,
                                                      const double* __restrict__ lsq,
                                                      double* __restrict__ G) {
    constexpr int U = D + 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    double qc[U], t[U][D];
#pragma unroll
    for (int k = 0; k < U; k++) {
        qc[k] = Q[(size_t)c * U + k];
#pragma unroll
        for (int d = 0; d < D; d++) t[k][d] = 0.0;
    }
    for (int j = 0; j < nslot; j++) {
        const int v = cf[(size_t)j * nc + c];
        if (v < 0) continue;
        const int f = v >> 1;
        const int nb = (v & 1) ? fc0[f] : fc1[f];
        if (nb < 0) continue;
        double g[D];
#pragma unroll
        for (int d = 0; d < D; d++) g[d] = lsq[((size_t)j * D + d) * nc + c];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const double dq = Q[(size_t)nb * U + k] - qc[k];
#pragma unroll
            for (int d = 0; d < D; d++) t[k][d] += g[d] * dq;
        }
    }
#pragma unroll
    for (int k = 0; k < U; k++)
#pragma unroll
        for (int d = 0; d < D; d++) G[((size_t)c * U + k) * D + d] = t[k][d];
}

// G[c][k][:] *= phi[c][k]: min / max over the cell and its face neighbours, slope tested at the
// face centres of the cell
template <int D>
__global__ void __launch_bounds__(256) k_limit(int nc, int nslot, int mode, const double* __restrict__ Q,
                                               const int32_t* __restrict__ cf,
                                               const int32_t* __restrict__ fc0,
                                               const int32_t* __restrict__ fc1,
                                               const double* __restrict__ dx0,
                                               const double* __restrict__ dx1,
                                               const double* __restrict__ eps2,
                                               double* __restrict__ G) {
    constexpr int U = D + 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    double qc[U], qmin[U], qmax[U], g[U][D];
    LimiterAcc acc[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
        qc[k] = qmin[k] = qmax[k] = Q[(size_t)c * U + k];
        acc[k].init(mode);
#pragma unroll
        for (int d = 0; d < D; d++) g[k][d] = G[((size_t)c * U + k) * D + d];
    }
    for (int j = 0; j < nslot; j++) {
        const int v = cf[(size_t)j * nc + c];
        if (v < 0) continue;
        const int f = v >> 1;
        const int nb = (v & 1) ? fc0[f] : fc1[f];
        if (nb < 0) continue;
#pragma unroll
        for (int k = 0; k < U; k++) {
            const double q = Q[(size_t)nb * U + k];
            qmin[k] = fmin(qmin[k], q);
            qmax[k] = fmax(qmax[k], q);
        }
    }
    const double e2 = eps2 ? eps2[c] : 0.0;
    for (int j = 0; j < nslot; j++) {
        const int v = cf[(size_t)j * nc + c];
        if (v < 0) continue;
        const int f = v >> 1;
        const double* dx = (v & 1) ? dx1 : dx0;
        double r[D];
#pragma unroll
        for (int d = 0; d < D; d++) r[d] = dx[(size_t)f * D + d];
#pragma unroll
        for (int k = 0; k < U; k++) {
            double dl = 0.0;
#pragma unroll
            for (int d = 0; d < D; d++) dl += g[k][d] * r[d];
            acc[k].add(mode, dl, qmax[k] - qc[k], qmin[k] - qc[k], e2);
        }
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
        const double phi = acc[k].phi(mode, qmax[k] - qc[k], qmin[k] - qc[k]);
#pragma unroll
        for (int d = 0; d < D; d++) G[((size_t)c * U + k) * D + d] = g[k][d] * phi;
    }
}

// min over the warp of POSITIVE doubles (their order is the order of their bit patterns)
__device__ __forceinline__ unsigned long long warp_min_bits(unsigned long long b) {
    const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}

// CFL step: dtmin = min_c V_c / sum_f (|u_c.S_f| + a_c |S_f|); bit pattern of a positive double
template <int D>
__global__ void __launch_bounds__(256) k_cfl(int n, int nc, int nslot, double gamma, const double* __restrict__ Q,
                                             const int32_t* __restrict__ cf,
                                             const double* __restrict__ Sd,
                                             const double* __restrict__ vol,
                                             unsigned long long* __restrict__ dtmin) {
    constexpr int U = D + 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bits = 0x7FF0000000000000ULL;  // +inf
    if (c < n) {
        double q[U];
#pragma unroll
        for (int k = 0; k < U; k++) q[k] = Q[(size_t)c * U + k];
        const double r = 1.0 / q[0];
        double m2 = 0.0;
#pragma unroll
        for (int d = 0; d < D; d++) m2 += q[d + 1] * q[d + 1];
        const double p = (q[U - 1] - 0.5 * m2 * r) * (gamma - 1.0);
        const double a = sqrt(gamma * p * r);
        double lam = 0.0;
        for (int j = 0; j < nslot; j++) {
            const int v = cf[(size_t)j * nc + c];
            if (v < 0) continue;
            const int f = v >> 1;
            double un = 0.0, s2 = 0.0;
#pragma unroll
            for (int d = 0; d < D; d++) {
                const double s = Sd[(size_t)f * D + d];
                un += q[d + 1] * r * s;
                s2 += s * s;
            }
            lam += fabs(un) + a * sqrt(s2);
        }
        const double t = vol[c] / lam;
        if (t > 0.0) bits = (unsigned long long)__double_as_longlong(t);  // NaN / non-positive never win
    }
    bits = warp_min_bits(bits);
    __shared__ unsigned long long sm[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long mbits = sm[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) mbits = min(mbits, sm[w]);
        atomicMin(dtmin, mbits);
    }
}

// dt = cfl * dtmin; time += dt; dtmin re-armed for the next step
__global__ void k_cfl_finish(double cfl, unsigned long long* dtmin, double* dt, double* time_acc) {
    const double t = cfl * __longlong_as_double((long long)*dtmin);
    *dt = t;
    *time_acc += t;
    *dtmin = 0x7FF0000000000000ULL;
}

// ---- implicit step (extension): the block system the LU-SGS sweeps solve ---------------------------
// A(Q,S) = d(F(Q).S)/dQ for a perfect gas, U x U row-major (same expressions as the oracle's fluxJacobian)
template <int D>
__device__ __forceinline__ void flux_jacobian(const double (&q)[D + 2], const double (&S)[D], double gamma, double (&A)[(D + 2) * (D + 2)]) {
    constexpr int U = D + 2;
    const double r = 1.0 / q[0];
    double u[D], un = 0.0, q2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; a++) { u[a] = q[a + 1] * r; un += u[a] * S[a]; q2 += u[a] * u[a]; }
    const double g1 = gamma - 1.0;
    const double phi = 0.5 * g1 * q2;
    const double p = (q[U - 1] - 0.5 * q[0] * q2) * g1;
    const double H = (q[U - 1] + p) * r;
#pragma unroll
    for (int i = 0; i < U * U; i++) A[i] = 0.0;
#pragma unroll
    for (int b = 0; b < D; b++) A[1 + b] = S[b];
#pragma unroll
    for (int a = 0; a < D; a++) {
        A[(1 + a) * U] = S[a] * phi - u[a] * un;
#pragma unroll
        for (int b = 0; b < D; b++) A[(1 + a) * U + 1 + b] = u[a] * S[b] - g1 * u[b] * S[a] + (a == b ? un : 0.0);
        A[(1 + a) * U + U - 1] = g1 * S[a];
    }
    A[(U - 1) * U] = (phi - H) * un;
#pragma unroll
    for (int b = 0; b < D; b++) A[(U - 1) * U + 1 + b] = H * S[b] - g1 * u[b] * un;
    A[(U - 1) * U + U - 1] = gamma * un;
}

template <int D>
__device__ __forceinline__ double spectral_radius(const double (&q)[D + 2], const double (&S)[D], double gamma) {
    constexpr int U = D + 2;
    const double r = 1.0 / q[0];
    double un = 0.0, q2 = 0.0, s2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; a++) { un += q[a + 1] * r * S[a]; q2 += q[a + 1] * q[a + 1]; s2 += S[a] * S[a]; }
    const double p = (q[U - 1] - 0.5 * q2 * r) * (gamma - 1.0);
    return fabs(un) + sqrt(gamma * p * r) * sqrt(s2);
}

// One thread per cell computes the diagonal block and the off-diagonal blocks of its row; each block is
// staged in shared memory and written by the whole CTA, so the 8*U*U-byte blocks leave as contiguous runs
// (a thread writing its own 200 bytes would touch 7 sectors with 25 separate stores).
template <int D>
__global__ void __launch_bounds__(128) k_assemble_implicit(int n, int nc, int nslot, double gamma, double dt,
                                                           const double* __restrict__ Q,
                                                           const int32_t* __restrict__ cf,
                                                           const int32_t* __restrict__ fc0,
                                                           const int32_t* __restrict__ fc1,
                                                           const double* __restrict__ Sd,
                                                           const double* __restrict__ vol,
                                                           const int32_t* __restrict__ dpos,
                                                           const int32_t* __restrict__ pos,
                                                           double* __restrict__ val) {
    constexpr int U = D + 2, UU = U * U, NT = 128;
    __shared__ double blk[NT][UU + 1];  // + 1: the rows of different threads start in different banks
    __shared__ int dst[NT];
    const int tid = threadIdx.x;
    const int c = blockIdx.x * NT + tid;
    const bool on = c < n;
    auto flush = [&]() {
        __syncthreads();
        for (int i = tid; i < NT * UU; i += NT) {
            const int t = i / UU, e = i - t * UU;
            if (dst[t] >= 0) val[(size_t)dst[t] * UU + e] = blk[t][e];
        }
        __syncthreads();
    };
    double qi[U], Dg[UU], A[UU];
#pragma unroll
    for (int k = 0; k < U; k++) qi[k] = on ? Q[(size_t)c * U + k] : 1.0;
#pragma unroll
    for (int i = 0; i < UU; i++) Dg[i] = 0.0;
    const double vdt = on ? vol[c] / dt : 0.0;
#pragma unroll
    for (int k = 0; k < U; k++) Dg[k * U + k] = vdt;
    for (int j = 0; j < nslot; j++) {
        const int v = on ? cf[(size_t)j * nc + c] : -1;
        int nb = -1;
        if (v >= 0) {
            const int f = v >> 1;
            const double sg = (v & 1) ? -1.0 : 1.0;
            nb = (v & 1) ? fc0[f] : fc1[f];
            double S[D];
#pragma unroll
            for (int d = 0; d < D; d++) S[d] = sg * Sd[(size_t)f * D + d];
            double lam = spectral_radius<D>(qi, S, gamma);
            double qj[U];
            if (nb >= 0) {
#pragma unroll
                for (int k = 0; k < U; k++) qj[k] = Q[(size_t)nb * U + k];
                lam = fmax(lam, spectral_radius<D>(qj, S, gamma));
            }
            flux_jacobian<D>(qi, S, gamma, A);
#pragma unroll
            for (int i = 0; i < UU; i++) Dg[i] += 0.5 * A[i];
#pragma unroll
            for (int k = 0; k < U; k++) Dg[k * U + k] += 0.5 * lam;
            if (nb >= 0) {
                flux_jacobian<D>(qj, S, gamma, A);
#pragma unroll
                for (int i = 0; i < UU; i++) {
                    double o = 0.5 * A[i];
                    if (i / U == i % U) o -= 0.5 * lam;
                    blk[tid][i] = o;
                }
            }
        }
        dst[tid] = nb >= 0 ? pos[(size_t)j * nc + c] : -1;
        flush();
    }
#pragma unroll
    for (int i = 0; i < UU; i++) blk[tid][i] = Dg[i];
    dst[tid] = on ? dpos[c] : -1;
    flush();
}

// Q_new = Q_old + dQ, with the residual of Time.cpp:69-76 and the NaN flag
template <int U>
__global__ void __launch_bounds__(256) k_add_increment(int n, const double* __restrict__ Qold, const double* __restrict__ dq,
                                                       double* __restrict__ Qnew,
                                                       unsigned long long* __restrict__ resid, int* __restrict__ nanflag) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    double r[U];
#pragma unroll
    for (int k = 0; k < U; k++) r[k] = 0.0;
    bool bad = false;
    if (c < n) {
#pragma unroll
        for (int k = 0; k < U; k++) {
            const double qo = Qold[(size_t)c * U + k];
            const double qn = qo + dq[(size_t)c * U + k];
            Qnew[(size_t)c * U + k] = qn;
            const double x = fabs(qn - qo) / qo;
            r[k] = (x > 0.0) ? x : 0.0;
            bad |= (qn != qn);
        }
    }
    __shared__ double sm[U][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < U; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r[k] = fmax(r[k], __shfl_xor_sync(0xffffffffu, r[k], o));
        if (lane == 0) sm[k][wid] = r[k];
    }
    const bool anybad = __syncthreads_or(bad);
    if (threadIdx.x < U) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmax(m, sm[threadIdx.x][w]);
        if (m > 0.0) atomicMax(&resid[threadIdx.x], (unsigned long long)__double_as_longlong(m));
    }
    if (anybad && threadIdx.x == 0) atomicOr(nanflag, 1);
}

__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

template <int D>
__global__ void __launch_bounds__(256) k_update(int nc, int n_upd, int nslot, int mode, double dt_val,
                                                const double* __restrict__ dt_dev,
                                                const double* __restrict__ Qold,
                                                const double* __restrict__ Phi,
                                                const int32_t* __restrict__ cf,
                                                const double* __restrict__ vol,
                                                double* __restrict__ Qnew,
                                                unsigned long long* __restrict__ resid,
                                                int* __restrict__ nanflag) {
    constexpr int U = D + 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const double dt = dt_dev ? *dt_dev : dt_val;  // device-resident dt: CFL stepping (extension)
    double r[U];
#pragma unroll
    for (int k = 0; k < U; k++) r[k] = 0.0;
    bool bad = false;
    if (c < n_upd) {
        double acc[U];
#pragma unroll
        for (int k = 0; k < U; k++) acc[k] = 0.0;
        for (int j = 0; j < nslot; j++) {
            const int v = cf[(size_t)j * nc + c];
            if (v < 0) continue;
            const int f = v >> 1;
            const double sg = (v & 1) ? -1.0 : 1.0;
#pragma unroll
            for (int k = 0; k < U; k++) acc[k] += sg * Phi[(size_t)f * U + k];
        }
        const double s = dt / vol[c];  // RhoSolver.cpp:64
#pragma unroll
        for (int k = 0; k < U; k++) {
            const double qo = Qold[(size_t)c * U + k];
            // mode 2: residual-vector mode (implicit step): -R_i instead of the update
            const double qn = (mode & 2) ? -acc[k] : qo - s * acc[k];
            Qnew[(size_t)c * U + k] = qn;
            // Time.cpp:72: signed denominator; NaN never wins, +inf can
            const double x = (mode & 2) ? 0.0 : fabs(qn - qo) / qo;
            r[k] = (x > 0.0) ? x : 0.0;
            bad |= (qn != qn);
        }
    }
    __shared__ double sm[U][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < U; k++) {
        const double m = warp_max(r[k]);
        if (lane == 0) sm[k][wid] = m;
    }
    const bool anybad = __syncthreads_or(bad);
    if (threadIdx.x < U) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmax(m, sm[threadIdx.x][w]);
        if (m > 0.0) atomicMax(&resid[threadIdx.x], (unsigned long long)__double_as_longlong(m));
    }
    if (anybad && threadIdx.x == 0) atomicOr(nanflag, 1);
}

// ---- peer-memory halo --------------------------------------------------------------------------------
// One kernel moves this rank's boundary rows straight into the ghost blocks of its neighbours (stores through
// peer pointers, NVLink), every thread fences at system scope, and the LAST CTA to finish publishes the epoch
// in each neighbour's flag array with a release store.  The receiver spins (bounded) on its own flags before
// the tiles that read ghost rows.  Why no "ready" handshake is needed for the ping-pong state: the rows of
// exchange e go into buffer Q[cur]; a neighbour last READ the ghost rows of that buffer two steps ago, and it
// cannot be more than one step behind (its step e-1 needed my exchange e-1).
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void k_peer_push(int n, int U, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                            const uint8_t* __restrict__ slot, const double* __restrict__ Q, PeerTable pt, int buf,
                            unsigned int* ticket, unsigned long long* epoch_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * U) {
        const int r = i / U, k = i - r * U;
        pt.q[buf][slot[r]][(size_t)dst[r] * U + k] = Q[(size_t)src[r] * U + k];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {  // every other CTA has fenced its stores before its ticket
            *ticket = 0;
            // the epoch lives on the device, so the exchange can sit in a CUDA graph that is launched many times
            const unsigned long long e = *epoch_dev + 1ULL;
            *epoch_dev = e;
            __threadfence_system();
            for (int j = 0; j < pt.nnb; j++) st_release_sys(pt.flag[j], e);
        }
    }
}

// bounded wait for the neighbours' next epoch: a lost peer must become an error, never a hung GPU
__global__ void k_peer_wait(const unsigned long long* flags, int nnb, unsigned long long* wait_dev, int* errflag) {
    const int j = threadIdx.x;
    const unsigned long long epoch = *wait_dev + 1ULL;
    // sticky: once a neighbour was lost, later exchanges return at once (the run ends with MSTGPU_ERR_NCCL at the
    // next residual / sync instead of spending the bound on every step)
    if (j < nnb && !(*(volatile int*)errflag & 4)) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + j) < epoch) {
            if (clock64() - t0 > 20000000000LL) { atomicOr(errflag, 4); break; }  // ~10 s
            __nanosleep(200);
        }
    }
    __syncwarp();
    if (j == 0) *wait_dev = epoch;
}

// halo: gather the rows a neighbour needs into a contiguous send buffer
__global__ void k_pack_rows(int n, int U, const int32_t* __restrict__ idx, const double* __restrict__ Q,
                            double* __restrict__ buf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * U) return;
    const int r = i / U, k = i - r * U;
    buf[i] = Q[(size_t)idx[r] * U + k];
}

// state in reference order <-> device order
__global__ void k_permute_in(int nc, int U, const double* __restrict__ src,
                             const int32_t* __restrict__ new2old, double* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nc * U) return;
    const int c = (int)(i / U), k = (int)(i % U);
    dst[i] = src[(size_t)new2old[c] * U + k];
}
__global__ void k_permute_out(int n, int W, const double* __restrict__ src,
                              const int32_t* __restrict__ new2old, double* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * W) return;
    const int c = (int)(i / W), k = (int)(i % W);
    dst[(size_t)new2old[c] * W + k] = src[i];
}

}  // namespace

// ============================================================================
// host side
// ============================================================================
namespace {

template <typename T>
int upload(mstgpu_ctx* ctx, T** dptr, const std::vector<T>& h) {
    const size_t bytes = h.size() * sizeof(T);
    CK(cudaMalloc((void**)dptr, bytes ? bytes : 8));
    ctx->dev_bytes += (int64_t)bytes;
    if (bytes) CK(cudaMemcpyAsync(*dptr, h.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    return MSTGPU_OK;
}
template <typename T>
int dalloc(mstgpu_ctx* ctx, T** dptr, size_t n) {
    CK(cudaMalloc((void**)dptr, n * sizeof(T)));
    ctx->dev_bytes += (int64_t)(n * sizeof(T));
    return MSTGPU_OK;
}

// per-face / per-cell tables of the split kernels (also read by the CFL reduction, the implicit assembly and the
// stage probes): uploaded from the plan's host copies on first use, which are dropped afterwards
int ensure_split_tables(mstgpu_ctx* ctx) {
    if (ctx->split_ready) return MSTGPU_OK;
    Plan& p = ctx->plan;
    int r;
    if ((r = upload(ctx, &ctx->fc0, p.fc0))) return r;
    if ((r = upload(ctx, &ctx->fc1, p.fc1))) return r;
    if ((r = upload(ctx, &ctx->Sd, p.Sd))) return r;
    if ((r = upload(ctx, &ctx->dx0, p.dx0))) return r;
    if ((r = upload(ctx, &ctx->dx1, p.dx1))) return r;
    if ((r = upload(ctx, &ctx->eta, p.eta))) return r;
    if ((r = upload(ctx, &ctx->meta, p.meta))) return r;
    if ((r = upload(ctx, &ctx->vol, p.vol))) return r;
    if ((r = upload(ctx, &ctx->cf, p.cf))) return r;
    if ((r = upload(ctx, &ctx->face_new2old, p.face_new2old))) return r;
    if (!p.lsq.empty() && (r = upload(ctx, &ctx->lsq, p.lsq))) return r;
    if (!p.eps2.empty() && (r = upload(ctx, &ctx->eps2, p.eps2))) return r;
    CK(cudaStreamSynchronize(ctx->stream));  // the uploads read pageable host vectors that are freed next
    p.fc0 = {}; p.fc1 = {}; p.Sd = {}; p.dx0 = {}; p.dx1 = {}; p.eta = {}; p.meta = {}; p.vol = {}; p.cf = {};
    p.lsq = {}; p.eps2 = {};
    ctx->split_ready = true;
    return MSTGPU_OK;
}

// the staging buffer holds at least n doubles
int ensure_stage(mstgpu_ctx* ctx, size_t n) {
    if (ctx->stage_cap >= n) return MSTGPU_OK;
    if (ctx->stage) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(ctx->stage));
        ctx->dev_bytes -= (int64_t)(ctx->stage_cap * sizeof(double));
        ctx->stage = nullptr; ctx->stage_cap = 0;
    }
    int r = dalloc(ctx, &ctx->stage, n);
    if (r) return r;
    ctx->stage_cap = n;
    return MSTGPU_OK;
}

struct KTimer {
    mstgpu_ctx* ctx;
    const char* name;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    // st: the stream the timed work is issued on (default: the compute stream); count = false: a span over
    // launches that are counted elsewhere
    KTimer(mstgpu_ctx* c, const char* n, cudaStream_t stream = nullptr, bool count = true) : ctx(c), name(n), st(stream ? stream : c->stream) {
        if (count) ctx->launches++;
        if (!ctx->ktiming) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!ctx->evpool.empty()) { e = ctx->evpool.back(); ctx->evpool.pop_back(); }
            else cudaEventCreate(&e);
            return e;
        };
        a = get(); b = get();
        cudaEventRecord(a, st);
    }
    ~KTimer() {
        if (!ctx->ktiming) { ctx->kstat[name].launches++; return; }
        cudaEventRecord(b, st);
        ctx->pending.push_back({name, {a, b}});
    }
};

void drain_timers(mstgpu_ctx* ctx) {
    for (auto& p : ctx->pending) {
        float ms = 0.f;
        cudaEventSynchronize(p.second.second);
        cudaEventElapsedTime(&ms, p.second.first, p.second.second);
        auto& s = ctx->kstat[p.first];
        s.ms += ms;
        s.launches++;
        ctx->evpool.push_back(p.second.first);
        ctx->evpool.push_back(p.second.second);
    }
    ctx->pending.clear();
}

int ensure_stage_buffers(mstgpu_ctx* ctx) {
    const size_t nq = (size_t)ctx->nc * ctx->U;
    if (!ctx->G) {
        int r;
        if ((r = dalloc(ctx, &ctx->G, nq * ctx->D))) return r;
        CK(cudaMemsetAsync(ctx->G, 0, nq * ctx->D * sizeof(double), ctx->stream));
    }
    if (ctx->cfg.viscous && !ctx->Gp) {
        int r;
        if ((r = dalloc(ctx, &ctx->Gp, (size_t)ctx->nc * (ctx->D + 1) * ctx->D))) return r;
    }
    if (!ctx->Phi) {
        int r;
        if ((r = dalloc(ctx, &ctx->Phi, (size_t)ctx->nf * ctx->U))) return r;
        CK(cudaMemsetAsync(ctx->Phi, 0, (size_t)ctx->nf * ctx->U * sizeof(double), ctx->stream));
    }
    return MSTGPU_OK;
}

#define NK(call)                                                                   \
    do {                                                                           \
        ncclResult_t r_ = (call);                                                  \
        if (r_ != ncclSuccess) {                                                   \
            set_error(ctx, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
            return MSTGPU_ERR_NCCL;                                                \
        }                                                                          \
    } while (0)

// ghost rows of Q <- owners' rows.  One pack kernel, one grouped send/recv; the
// receives land directly in the ghost block of Q (ghosts are ordered by owner).
int halo_exchange(mstgpu_ctx* ctx, double* Q, cudaStream_t st) {
    if (!ctx->partitioned || ctx->halo.empty()) return MSTGPU_OK;
    // timing diagnosis only (WRONG results: the ghost rows keep their old values): every rank runs its tiles
    // uncoupled from its neighbours, which separates a slow GPU / partition from waiting on the exchange
    static const bool no_halo = getenv("MSTGPU_DEBUG_NO_HALO") != nullptr;
    if (no_halo) return MSTGPU_OK;
    if (!ctx->comm) { set_error(ctx, "partitioned context without a communicator: call mstgpu_comm_init"); return MSTGPU_ERR_STATE; }
    const int U = ctx->U;
    if (ctx->peer_ok && (Q == ctx->Q[0] || Q == ctx->Q[1])) {
        // boundary rows -> the neighbours' ghost blocks of the same buffer, then the epoch; wait for theirs
        const int buf = Q == ctx->Q[0] ? 0 : 1;
        const int n = std::max(1, ctx->send_total * U);
        k_peer_push<<<(n + 255) / 256, 256, 0, st>>>(ctx->send_total, U, ctx->send_idx, ctx->push_dst, ctx->push_slot, Q, ctx->peer, buf,
                                                     ctx->push_ticket, ctx->peer_epochs);
        k_peer_wait<<<1, 32, 0, st>>>(ctx->peer_flags, ctx->peer.nnb, ctx->peer_epochs + 1, ctx->nanflag);
        ctx->launches += 2;
        return MSTGPU_OK;
    }
    if (ctx->send_total > 0) {
        ctx->launches++;
        k_pack_rows<<<(ctx->send_total * U + 255) / 256, 256, 0, st>>>(ctx->send_total, U, ctx->send_idx, Q, ctx->sendbuf);
    }
    NK(g_nccl.GroupStart());
    for (const auto& h : ctx->halo) {
        if (h.send_count) NK(g_nccl.Send(ctx->sendbuf + (size_t)h.send_off * U, (size_t)h.send_count * U, ncclDouble, h.rank, ctx->comm, st));
        if (h.recv_count) NK(g_nccl.Recv(Q + (size_t)h.recv_first * U, (size_t)h.recv_count * U, ncclDouble, h.rank, ctx->comm, st));
    }
    NK(g_nccl.GroupEnd());
    return MSTGPU_OK;
}

template <int D, int ORDER, int NT, int NS, bool LIM, bool VISC, int VAR>
int launch_tiles_var(mstgpu_ctx* ctx, double dt, const double* dtd, const double* Qo, double* Qn, int want_resid, int which, cudaStream_t st) {
    auto kern = k_step_tiles<D, ORDER, NT, NS, LIM, VISC, VAR>;
    // the opt-in to > 48 KB of dynamic shared memory is a per-DEVICE attribute of the function: remembered per
    // context (a context lives on one device), not per thread -- one host thread may drive several devices
    size_t& configured_smem = ctx->smem_configured[(const void*)kern];
    if (configured_smem < ctx->tile_smem) {
        // never lower it: another context on this device may have opted in to more for the same instantiation
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, kern));
        if ((size_t)fa.maxDynamicSharedSizeBytes < ctx->tile_smem)
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->tile_smem));
        configured_smem = ctx->tile_smem;
    }
    // which: 0 = tiles without ghost cells, 1 = tiles whose rings hold ghost cells, 2 = all
    int ci = 0, cidx = -1;
    for (const auto& tc : ctx->tile_classes) {
        cidx++;
        if (which != 2 && (int)tc.halo != which) continue;
        if (ctx->sub_group >= 0) {
            // streamed step: the tiles of this class whose input is complete with host chunk sub_group
            const std::vector<int>& off = ctx->hs.off[cidx];
            const int first = off[ctx->sub_group], cnt = off[ctx->sub_group + 1] - first;
            if (cnt <= 0) continue;
            TileArrays ta = ctx->ta;
            ta.order = ctx->hs.d_order;
            kern<<<cnt, NT, tc.smem, st>>>(ta, first, cnt, 0, want_resid, ctx->dcfg, dt, dtd, Qo, Qn, ctx->resid, ctx->nanflag);
            ctx->launches++;
            continue;
        }
        cudaStream_t s = (ctx->fork_stream && (ci++ & 1)) ? ctx->fork_stream : st;
        int grid = tc.count, wave = 0;
        if (VAR & 3) {  // experimental variants: CTAs resident at once for this class's shared-memory size
            int per_sm = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, tc.smem));
            wave = std::max(1, per_sm) * ctx->sm_count;
            if (VAR & 2) grid = std::min(tc.count, wave);
        }
        kern<<<grid, NT, tc.smem, s>>>(ctx->ta, tc.first, tc.count, wave, want_resid, ctx->dcfg, dt, dtd, Qo, Qn, ctx->resid, ctx->nanflag);
        ctx->launches++;
    }
    return MSTGPU_OK;
}

template <int D, int ORDER, int NT, int NS, bool LIM, bool VISC>
int launch_tiles_lim(mstgpu_ctx* ctx, double dt, const double* dtd, const double* Qo, double* Qn, int want_resid, int which, cudaStream_t st) {
    return launch_tiles_var<D, ORDER, NT, NS, LIM, VISC, 0>(ctx, dt, dtd, Qo, Qn, want_resid, which, st);
}

template <int D, int ORDER, int NT, int NS>
int launch_tiles(mstgpu_ctx* ctx, double dt, const double* dtd, const double* Qo, double* Qn, int wr, int which, cudaStream_t st) {
    if (ORDER == 2) {
        const bool lim = ctx->cfg.limiter != 0, visc = ctx->cfg.viscous != 0;
        if (lim && visc) return launch_tiles_lim<D, 2, NT, NS, true, true>(ctx, dt, dtd, Qo, Qn, wr, which, st);
        if (lim) return launch_tiles_lim<D, 2, NT, NS, true, false>(ctx, dt, dtd, Qo, Qn, wr, which, st);
        if (visc) return launch_tiles_lim<D, 2, NT, NS, false, true>(ctx, dt, dtd, Qo, Qn, wr, which, st);
    }
    if (D == 3 && ORDER == 2 && NT == 256 && NS == 5 && ctx->tile_var) {
        // experimental variants of the default instantiation (step_tiles.cuh), bit-identical results
        switch (ctx->tile_var) {
            case 1: return launch_tiles_var<3, 2, 256, 5, false, false, 1>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            case 2: return launch_tiles_var<3, 2, 256, 5, false, false, 2>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            case 3: return launch_tiles_var<3, 2, 256, 5, false, false, 3>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            case 4: return launch_tiles_var<3, 2, 256, 5, false, false, 4>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            case 5: return launch_tiles_var<3, 2, 256, 5, false, false, 5>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            case 7: return launch_tiles_var<3, 2, 256, 5, false, false, 7>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            default: break;
        }
    }
    if (D == 3 && ORDER == 2 && NT == 128 && NS == 5 && !(ctx->tile_var & 32)) {
        // 128-thread CTAs (the default for tets at second order): registers sized for 4 resident CTAs per SM
        // (MSTGPU_TILE_VAR 64: for 5, experimental; 32: the plain 3-CTA allocation)
        // experiments on the packet stream of phase 2 (step_tiles.cuh): 256 first face's words before the state wait,
        // 512 next trip's lines into L1, 1024 next face's words requested after the reconstruction
        switch (ctx->tile_var & (256 | 512 | 1024)) {
            case 256: return launch_tiles_var<3, 2, 128, 5, false, false, 8 | 256>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            case 512: return launch_tiles_var<3, 2, 128, 5, false, false, 8 | 512>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            case 256 | 512: return launch_tiles_var<3, 2, 128, 5, false, false, 8 | 256 | 512>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            case 1024: return launch_tiles_var<3, 2, 128, 5, false, false, 8 | 1024>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            default: break;
        }
        return (ctx->tile_var & 64) ? launch_tiles_var<3, 2, 128, 5, false, false, 64>(ctx, dt, dtd, Qo, Qn, wr, which, st)
                                    : launch_tiles_var<3, 2, 128, 5, false, false, 8>(ctx, dt, dtd, Qo, Qn, wr, which, st);
    }
    if (D == 2 && ORDER == 1 && NT == 256 && NS == 4 && !(ctx->tile_var & 32)) {
        // 2-D first order on triangles is occupancy-bound (light flux, ~10 us of dependent latency per tile):
        // the register allocation is sized for 4 (AUSM+) / 3 (Roe) resident CTAs per SM instead of 2.
        // Measured on 998 046 triangles (profiles/r1c_ab_variants.json): AUSM+ 78.0 -> 60.6 us per step,
        // Roe 57.4 -> 47.2 us; bit-identical results.  MSTGPU_TILE_VAR: 8 / 16 force one of the two, 32 = plain.
        const bool four = (ctx->tile_var & 8) || (!(ctx->tile_var & 16) && ctx->cfg.flux == MSTGPU_FLUX_AUSM);
        return four ? launch_tiles_var<2, 1, 256, 4, false, false, 8>(ctx, dt, dtd, Qo, Qn, wr, which, st)
                    : launch_tiles_var<2, 1, 256, 4, false, false, 16>(ctx, dt, dtd, Qo, Qn, wr, which, st);
    }
    return launch_tiles_lim<D, ORDER, NT, NS, false, false>(ctx, dt, dtd, Qo, Qn, wr, which, st);
}

template <int D, int NS>
int launch_tiles_ns(mstgpu_ctx* ctx, double dt, const double* dtd, const double* Qo, double* Qn, int wr, int which, cudaStream_t st) {
    const bool o2 = ctx->cfg.order == 2;
    if constexpr (D == 3 && NS == 5) {
        if (ctx->tile_staged && o2) {
            if (ctx->tile_NT == 128) return launch_tiles_var<3, 2, 128, 5, false, false, 128 + 8>(ctx, dt, dtd, Qo, Qn, wr, which, st);
            return launch_tiles_var<3, 2, 256, 5, false, false, 128>(ctx, dt, dtd, Qo, Qn, wr, which, st);
        }
        // wider CTAs for the default scheme (tets, second order, no extension): 2 x 320 threads at 96 registers
        // or 2 x 384 threads at 80 registers per SM instead of 2 x 256 at 128
        if (o2 && !ctx->cfg.limiter && !ctx->cfg.viscous && ctx->tile_NT == 320)
            return launch_tiles_var<3, 2, 320, 5, false, false, 0>(ctx, dt, dtd, Qo, Qn, wr, which, st);
        if (o2 && !ctx->cfg.limiter && !ctx->cfg.viscous && ctx->tile_NT == 384)
            return launch_tiles_var<3, 2, 384, 5, false, false, 0>(ctx, dt, dtd, Qo, Qn, wr, which, st);
    }
    if (ctx->tile_NT == 128) return o2 ? launch_tiles<D, 2, 128, NS>(ctx, dt, dtd, Qo, Qn, wr, which, st) : launch_tiles<D, 1, 128, NS>(ctx, dt, dtd, Qo, Qn, wr, which, st);
    return o2 ? launch_tiles<D, 2, 256, NS>(ctx, dt, dtd, Qo, Qn, wr, which, st) : launch_tiles<D, 1, 256, NS>(ctx, dt, dtd, Qo, Qn, wr, which, st);
}

template <int D>
int launch_tiles_any(mstgpu_ctx* ctx, double dt, const double* dtd, const double* Qo, double* Qn, int wr, int which, cudaStream_t st) {
    // stencil size = 1 + faces per cell: triangles 4, tets / quads 5, hexes 7
    switch (ctx->nslot) {
        case 3: return launch_tiles_ns<D, 4>(ctx, dt, dtd, Qo, Qn, wr, which, st);
        case 4: return launch_tiles_ns<D, 5>(ctx, dt, dtd, Qo, Qn, wr, which, st);
        case 6: return launch_tiles_ns<D, 7>(ctx, dt, dtd, Qo, Qn, wr, which, st);
        default: set_error(ctx, "fused kernel supports cells with 3, 4 or 6 faces"); return MSTGPU_ERR_ARG;
    }
}

// dt of the step about to run, computed on the device from Q (extension, mstgpu_step_cfl):
// cell minimum -> [min over ranks] -> dt_dev[0] = cfl * min, dt_dev[1] += dt
template <int D>
int cfl_on_device(mstgpu_ctx* ctx, double cfl, const double* Q) {
    { int r0 = ensure_split_tables(ctx); if (r0) return r0; }
    k_cfl<D><<<(ctx->n_owned + 255) / 256, 256, 0, ctx->stream>>>(ctx->n_owned, ctx->nc, ctx->nslot, ctx->dcfg.gamma, Q, ctx->cf,
                                                                 ctx->Sd, ctx->vol, ctx->dtmin);
    if (ctx->comm) NK(g_nccl.AllReduce(ctx->dtmin, ctx->dtmin, 1, ncclUint64, ncclMin, ctx->comm, ctx->stream));
    k_cfl_finish<<<1, 1, 0, ctx->stream>>>(cfl, ctx->dtmin, ctx->dt_dev, ctx->dt_dev + 1);
    ctx->launches += 2;
    return MSTGPU_OK;
}

// One step of the fused path: Qc -> Qn.  Without neighbours: every tile class on the compute stream (under graph
// capture the classes fork onto the second stream).  With neighbours: the exchange and the tiles whose rings
// hold ghost cells on the high-priority second stream, the interior tiles on the compute stream meanwhile.
template <int D>
int issue_tile_step(mstgpu_ctx* ctx, double dt, const double* dtd, double* Qc, double* Qn, int wr, bool overlap, bool capturing) {
    int r;
    if (overlap) {
        // comm stream: ghost rows <- owners, as soon as the previous step is complete;
        // compute stream: tiles that touch no ghost cell meanwhile, the others after
        // (the halo tiles follow the exchange on the comm stream, so they fill the SMs
        // the interior launch leaves idle in its tail instead of waiting behind it)
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_done, 0));
        {   // per-step breakdown when kernel timing is on (bench.py): exchange, halo tiles, interior tiles
            KTimer t(ctx, "halo_exchange", ctx->stream2, false);
            if ((r = halo_exchange(ctx, Qc, ctx->stream2))) return r;
        }
        {
            KTimer t(ctx, "halo_tiles", ctx->stream2, false);
            if ((r = launch_tiles_any<D>(ctx, dt, dtd, Qc, Qn, wr, 1, ctx->stream2))) return r;
        }
        CK(cudaEventRecord(ctx->ev_halo, ctx->stream2));
        {
            KTimer t(ctx, "interior_tiles", ctx->stream, false);
            if ((r = launch_tiles_any<D>(ctx, dt, dtd, Qc, Qn, wr, 0, ctx->stream))) return r;
        }
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
        CK(cudaEventRecord(ctx->ev_done, ctx->stream));
        return MSTGPU_OK;
    }
    const bool fork = capturing && ctx->tile_classes.size() > 1;
    if (fork) {
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        ctx->fork_stream = ctx->stream2;
    }
    r = launch_tiles_any<D>(ctx, dt, dtd, Qc, Qn, wr, 2, ctx->stream);
    ctx->fork_stream = nullptr;
    if (fork) {
        CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    return r;
}

// Two fixed-dt steps (Q[cur] -> Q[cur^1] -> Q[cur]) captured once as a CUDA graph: the host issues
// one launch per pair of steps, and the tile classes of a step (separate launches because their
// shared-memory sizes differ) run as parallel branches instead of one behind the other's tail.
// A partitioned step is captured too when its halo goes through peer memory (kernels only: the push, the
// bounded flag wait, the two tile classes on their two streams; the epochs live on the device).
template <int D>
int build_step_graph(mstgpu_ctx* ctx, double dt, int start_cur, bool overlap) {
    cudaGraph_t g = nullptr;
    // which = 3 selects no tile class: only the kernel's shared-memory attribute is set, outside the capture
    int r = launch_tiles_any<D>(ctx, dt, nullptr, ctx->Q[0], ctx->Q[1], 0, 3, ctx->stream);
    if (r) return r;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int64_t before = ctx->launches;
    if (overlap) cudaEventRecord(ctx->ev_done, ctx->stream);
    for (int s = 0; s < 2 && r == MSTGPU_OK; s++)
        r = issue_tile_step<D>(ctx, dt, nullptr, ctx->Q[start_cur ^ s], ctx->Q[start_cur ^ s ^ 1], 0, overlap, true);
    ctx->graph_launches = ctx->launches - before;
    ctx->launches = before;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    if (r != MSTGPU_OK) { if (g) cudaGraphDestroy(g); return r; }
    if (e != cudaSuccess) { set_error(ctx, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); return MSTGPU_ERR_CUDA; }
    e = cudaGraphInstantiate(&ctx->step_graph[start_cur], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { set_error(ctx, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); return MSTGPU_ERR_CUDA; }
    return MSTGPU_OK;
}

// Will a call of nsteps fixed-dt steps go through the CUDA graph?  Builds the executable for the current buffer on
// first use (and after a change of dt), so that callers that time a call can keep the one-time build out of it.
// The graph pays where a step is launch-bound (SOD tube: 15.6 -> 9.3 us per step, 1 M triangles: 65 -> 60 us).  On
// large meshes direct launches are the faster path -- measured at 50.2 M tets: 5.43 against 5.50 ms per step on one
// GPU, 0.790 against 0.833 ms on eight (profiles/r2_scaling.md) -- so the graph is used up to 8192 tiles
// (MSTGPU_GRAPH=1 forces it, MSTGPU_NO_GRAPH disables it).  With neighbours only when the halo is peer memory (no
// NCCL call inside the capture).
template <int D>
bool step_graph_ready(mstgpu_ctx* ctx, double dt, int nsteps, double cfl) {
    const bool overlap = ctx->partitioned && !ctx->halo.empty();
    static const bool no_graph = getenv("MSTGPU_NO_GRAPH") != nullptr;
    static const bool force_graph = getenv("MSTGPU_GRAPH") != nullptr;
    if (!ctx->use_tiles || !((!overlap || ctx->peer_ok) && cfl <= 0.0 && !ctx->ktiming && !no_graph && nsteps >= 4 && (ctx->ntiles <= 8192 || force_graph)))
        return false;
    if (overlap && !ctx->comm) return false;
    if (ctx->step_graph_dt != dt) {
        for (auto& ge : ctx->step_graph) if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
        ctx->step_graph_dt = dt;
    }
    if (!ctx->step_graph[ctx->cur]) ctx->graph_rc = build_step_graph<D>(ctx, dt, ctx->cur, overlap);  // failure leaves it null
    return true;
}

template <int D>
int step_tiles_impl(mstgpu_ctx* ctx, double dt, int nsteps, double cfl) {
    const double* dtd = cfl > 0.0 ? ctx->dt_dev : nullptr;
    const bool overlap = ctx->partitioned && !ctx->halo.empty();
    if (overlap && !ctx->comm && !getenv("MSTGPU_DEBUG_NO_HALO")) { set_error(ctx, "partitioned context without a communicator: call mstgpu_comm_init"); return MSTGPU_ERR_STATE; }
    // long fixed-dt runs: pairs of steps from the graph, the last one or two steps (the observable residual
    // belongs to the last) launched directly.  With neighbours only when the halo is peer memory (no NCCL call
    // inside the capture).
    if (step_graph_ready<D>(ctx, dt, nsteps, cfl)) {
        if (!ctx->step_graph[ctx->cur]) return ctx->graph_rc ? ctx->graph_rc : MSTGPU_ERR_CUDA;  // the build failed; its message is in ctx->err
        const int pairs = (nsteps - 1) / 2;
        for (int i = 0; i < pairs; i++) CK(cudaGraphLaunch(ctx->step_graph[ctx->cur], ctx->stream));
        ctx->launches += pairs * ctx->graph_launches;
        nsteps -= 2 * pairs;  // cur is unchanged after an even number of steps
        ctx->stepped = true;
        ctx->probes_valid = false;
    }
    if (overlap && nsteps > 0) CK(cudaEventRecord(ctx->ev_done, ctx->stream));
    for (int s = 0; s < nsteps; s++) {
        double* Qc = ctx->Q[ctx->cur];
        double* Qn = ctx->Q[ctx->cur ^ 1];
        // the residual of a step is observable only for the last step of the call
        const int wr = (s == nsteps - 1) ? 1 : 0;
        if (wr) CK(cudaMemsetAsync(ctx->resid, 0, 8 * sizeof(unsigned long long), ctx->stream));
        int r;
        if (cfl > 0.0) {
            if ((r = cfl_on_device<D>(ctx, cfl, Qc))) return r;
            if (overlap) CK(cudaEventRecord(ctx->ev_done, ctx->stream));  // the comm stream needs dt too
        }
        KTimer t(ctx, "step_tiles");
        ctx->launches--;  // the launches are counted one by one in launch_tiles
        if ((r = issue_tile_step<D>(ctx, dt, dtd, Qc, Qn, wr, overlap, false))) return r;
        ctx->cur ^= 1;
    }
    CK(cudaGetLastError());
    ctx->stepped = nsteps > 0 || ctx->stepped;
    if (nsteps > 0) ctx->probes_valid = false;
    return MSTGPU_OK;
}

// the gradient stage of the split path: Green-Gauss (the reference) or least squares, then the limiter
template <int D>
void launch_gradient_stage(mstgpu_ctx* ctx, const double* Qo) {
    const int nc = ctx->nc;
    const bool lsq = ctx->cfg.gradient == MSTGPU_GRAD_LSQ && ctx->cfg.order == 2;
    if ((ctx->cfg.order == 2 && !lsq) || ctx->cfg.viscous) {
        KTimer t(ctx, "gradient");
        k_gradient<D><<<(nc + 255) / 256, 256, 0, ctx->stream>>>(
            nc, ctx->nslot, Qo, ctx->cf, ctx->fc0, ctx->fc1, ctx->eta, ctx->Sd, ctx->vol, ctx->G, ctx->Gp, ctx->dcfg.cv);
    }
    if (lsq) {
        KTimer t(ctx, "gradient_lsq");
        k_gradient_lsq<D><<<(nc + 255) / 256, 256, 0, ctx->stream>>>(nc, ctx->nslot, Qo, ctx->cf, ctx->fc0, ctx->fc1, ctx->lsq, ctx->G);
    }
    if (ctx->cfg.order == 2 && ctx->cfg.limiter != 0) {
        KTimer t(ctx, "limiter");
        k_limit<D><<<(nc + 255) / 256, 256, 0, ctx->stream>>>(nc, ctx->nslot, ctx->cfg.limiter, Qo, ctx->cf, ctx->fc0, ctx->fc1,
                                                              ctx->dx0, ctx->dx1, ctx->eps2, ctx->G);
    }
}

template <int D>
int step_impl(mstgpu_ctx* ctx, double dt, int nsteps, double cfl = 0.0) {
    if (ctx->use_tiles) return step_tiles_impl<D>(ctx, dt, nsteps, cfl);
    { int r0 = ensure_split_tables(ctx); if (r0) return r0; }
    const double* dtd = cfl > 0.0 ? ctx->dt_dev : nullptr;
    const int nc = ctx->nc, nf = ctx->nf;
    {
        int r = ensure_stage_buffers(ctx);
        if (r) return r;
    }
    if (nsteps > 0) ctx->probes_valid = true;
    for (int s = 0; s < nsteps; s++) {
        {
            int r = halo_exchange(ctx, ctx->Q[ctx->cur], ctx->stream);
            if (r) return r;
        }
        const double* Qo = ctx->Q[ctx->cur];
        double* Qn = ctx->Q[ctx->cur ^ 1];
        CK(cudaMemsetAsync(ctx->resid, 0, 8 * sizeof(unsigned long long), ctx->stream));
        if (cfl > 0.0) {
            int r = cfl_on_device<D>(ctx, cfl, Qo);
            if (r) return r;
        }
        launch_gradient_stage<D>(ctx, Qo);
        {
            KTimer t(ctx, "flux");
            if (ctx->cfg.order == 2)
                k_flux<D, 2><<<(nf + 127) / 128, 128, 0, ctx->stream>>>(
                    nf, ctx->dcfg, Qo, ctx->G, ctx->fc0, ctx->fc1, ctx->meta, ctx->Sd, ctx->dx0, ctx->dx1, ctx->Phi, ctx->Gp, ctx->eta);
            else
                k_flux<D, 1><<<(nf + 127) / 128, 128, 0, ctx->stream>>>(
                    nf, ctx->dcfg, Qo, ctx->G, ctx->fc0, ctx->fc1, ctx->meta, ctx->Sd, ctx->dx0, ctx->dx1, ctx->Phi, ctx->Gp, ctx->eta);
        }
        {
            KTimer t(ctx, "update");
            k_update<D><<<(ctx->n_owned + 255) / 256, 256, 0, ctx->stream>>>(
                nc, ctx->n_owned, ctx->nslot, 1, dt, dtd, Qo, ctx->Phi, ctx->cf, ctx->vol, Qn, ctx->resid, ctx->nanflag);
        }
        ctx->cur ^= 1;  // RhoSolver::updateNewToOld as a pointer swap
    }
    CK(cudaGetLastError());
    ctx->stepped = nsteps > 0 || ctx->stepped;
    return MSTGPU_OK;
}

// Under the fused kernel the stages never reach HBM; recompute them from the
// state the last step started from (Q[cur^1]) with the split kernels.
template <int D>
int recompute_stages(mstgpu_ctx* ctx) {
    int r = ensure_split_tables(ctx);
    if (r) return r;
    if ((r = ensure_stage_buffers(ctx))) return r;
    const int nc = ctx->nc, nf = ctx->nf;
    const double* Qo = ctx->Q[ctx->cur ^ 1];
    launch_gradient_stage<D>(ctx, Qo);
    if (ctx->cfg.order == 2)
        k_flux<D, 2><<<(nf + 127) / 128, 128, 0, ctx->stream>>>(nf, ctx->dcfg, Qo, ctx->G, ctx->fc0, ctx->fc1, ctx->meta,
                                                                 ctx->Sd, ctx->dx0, ctx->dx1, ctx->Phi, ctx->Gp, ctx->eta);
    else
        k_flux<D, 1><<<(nf + 127) / 128, 128, 0, ctx->stream>>>(nf, ctx->dcfg, Qo, ctx->G, ctx->fc0, ctx->fc1, ctx->meta,
                                                                 ctx->Sd, ctx->dx0, ctx->dx1, ctx->Phi, ctx->Gp, ctx->eta);
    ctx->launches += 1;
    CK(cudaGetLastError());
    ctx->probes_valid = true;
    return MSTGPU_OK;
}

int fetch_permuted(mstgpu_ctx* ctx, const double* dsrc, const int32_t* new2old, int n, int W, double* host) {
    const size_t tot = (size_t)n * W;
    { int r0 = ensure_stage(ctx, tot); if (r0) return r0; }
    k_permute_out<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(n, W, dsrc, new2old, ctx->stage);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host, ctx->stage, tot * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MSTGPU_OK;
}

// ---- streamed step (mstgpu_step_host) ------------------------------------------------------------
// rows [r0, r0 + n) of the staging buffer (reference order) -> their device-order rows of Q, and back
__global__ void k_rows_in(int64_t r0, int64_t n, int U, const double* __restrict__ stage,
                          const int32_t* __restrict__ old2new, double* __restrict__ Q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * U) return;
    const int64_t r = r0 + i / U;
    const int k = (int)(i % U);
    Q[(size_t)old2new[r] * U + k] = stage[(size_t)r * U + k];
}
__global__ void k_rows_out(int64_t r0, int64_t n, int U, const double* __restrict__ Q,
                           const int32_t* __restrict__ old2new, double* __restrict__ stage) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * U) return;
    const int64_t r = r0 + i / U;
    const int k = (int)(i % U);
    stage[(size_t)r * U + k] = Q[(size_t)old2new[r] * U + k];
}

void host_stream_free(mstgpu_ctx* ctx) {
    auto& h = ctx->hs;
    for (auto e : h.ev_in) cudaEventDestroy(e);
    for (auto e : h.ev_out) cudaEventDestroy(e);
    if (h.ev_start) cudaEventDestroy(h.ev_start);
    if (h.s_in) cudaStreamDestroy(h.s_in);
    if (h.s_out) cudaStreamDestroy(h.s_out);
    if (h.d_order) { cudaFree(h.d_order); ctx->dev_bytes -= (int64_t)ctx->ntiles * 4; }
    if (h.d_old2new) { cudaFree(h.d_old2new); ctx->dev_bytes -= (int64_t)ctx->n_owned * 4; }
    h = mstgpu_ctx::HostStream{};
}

// Schedule of the streamed step for `nchunks` chunks of host rows:
//   group(tile)  = the chunk whose arrival completes the tile's input (tiles with ghost cells in a ring: after
//                  the halo exchange, group nchunks); the launch list of every tile class is sorted by group
//   done(chunk)  = the last group that writes one of the chunk's rows: the chunk can leave after it
// Whatever the host numbering is, the schedule is valid; how much of the copies it hides depends on how well
// the host order follows the mesh (a mesher's cell order usually does; a random one degenerates to copy in,
// step, copy out).
int host_stream_setup(mstgpu_ctx* ctx, int nchunks) {
    auto& h = ctx->hs;
    const int64_t n = ctx->n_owned;
    nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(nchunks, std::max<int64_t>(1, n / 1024)));
    if (h.nchunks == nchunks && h.d_order) return MSTGPU_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    host_stream_free(ctx);
    const int64_t rows = (n + nchunks - 1) / nchunks;
    nchunks = (int)((n + rows - 1) / rows);
    const int nt = ctx->ntiles, G = nchunks;
    std::vector<int> grp(nt);
    std::vector<int32_t> order(nt);
    h.off.assign(ctx->tile_classes.size(), std::vector<int>(G + 2, 0));
    for (size_t ci = 0; ci < ctx->tile_classes.size(); ci++) {
        const auto& tc = ctx->tile_classes[ci];
        std::vector<int> cnt(G + 2, 0);
        for (int t = tc.first; t < tc.first + tc.count; t++) {
            grp[t] = tc.halo ? G : (int)(ctx->tile_ready_row[t] / rows);
            cnt[grp[t] + 1]++;
        }
        auto& off = h.off[ci];
        off[0] = tc.first;
        for (int g = 0; g <= G; g++) off[g + 1] = off[g] + cnt[g + 1];
        std::vector<int> pos(off.begin(), off.end() - 1);
        for (int t = tc.first; t < tc.first + tc.count; t++) order[pos[grp[t]]++] = t;
    }
    // the group after which every row of a host chunk has been written
    std::vector<int> done(G, 0);
    const std::vector<int32_t>& n2o = ctx->plan.cell_new2old;
#pragma omp parallel
    {
        std::vector<int> mine(G, 0);
#pragma omp for schedule(static) nowait
        for (int t = 0; t < nt; t++)
            for (int c = ctx->tile_cb_h[t]; c < ctx->tile_cb_h[t] + ctx->tile_nown_h[t]; c++) {
                const int j = (int)(n2o[c] / rows);
                mine[j] = std::max(mine[j], grp[t]);
            }
#pragma omp critical
        for (int j = 0; j < G; j++) done[j] = std::max(done[j], mine[j]);
    }
    h.out_chunks.assign(G + 1, {});
    for (int j = 0; j < G; j++) h.out_chunks[done[j]].push_back(j);
    int r;
    if ((r = upload(ctx, &h.d_order, order))) return r;
    std::vector<int32_t> o2n(ctx->plan.cell_old2new.begin(), ctx->plan.cell_old2new.begin() + n);
    if ((r = upload(ctx, &h.d_old2new, o2n))) return r;
    CK(cudaStreamSynchronize(ctx->stream));  // the uploads read the vectors above
    CK(cudaStreamCreateWithFlags(&h.s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h.s_out, cudaStreamNonBlocking));
    h.ev_in.resize(G); h.ev_out.resize(G + 1);
    for (auto& e : h.ev_in) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : h.ev_out) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h.ev_start, cudaEventDisableTiming));
    h.nchunks = G;
    h.chunk_rows = rows;
    return MSTGPU_OK;
}

template <int D>
int step_host_impl(mstgpu_ctx* ctx, const double* qin, double* qout, double dt) {
    auto& h = ctx->hs;
    const int U = ctx->U, G = h.nchunks;
    const int64_t n = ctx->n_owned, rows = h.chunk_rows;
    double* Qc = ctx->Q[ctx->cur];
    double* Qn = ctx->Q[ctx->cur ^ 1];
    const bool halo = ctx->partitioned && !ctx->halo.empty();
    int r = launch_tiles_any<D>(ctx, dt, nullptr, Qc, Qn, 0, 3, ctx->stream);  // which = 3: shared-memory opt-in only
    if (r) return r;
    CK(cudaMemsetAsync(ctx->resid, 0, 8 * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->nanflag, 0, sizeof(int), ctx->stream));
    CK(cudaEventRecord(h.ev_start, ctx->stream));  // the staging buffer is free once earlier work has drained
    CK(cudaStreamWaitEvent(h.s_in, h.ev_start, 0));
    auto span = [&](int j, int64_t& r0, int64_t& cnt) { r0 = (int64_t)j * rows; cnt = std::min(rows, n - r0); };
    // rows that are final after group g: device order -> staging buffer (the chunk's input was consumed long
    // ago) on the compute stream, then device -> host on the output stream
    auto emit = [&](int g) -> int {
        const std::vector<int>& ch = h.out_chunks[g];
        if (ch.empty()) return MSTGPU_OK;
        for (int j : ch) {
            int64_t r0, cnt; span(j, r0, cnt);
            k_rows_out<<<(unsigned)((cnt * U + 255) / 256), 256, 0, ctx->stream>>>(r0, cnt, U, Qn, h.d_old2new, ctx->stage);
            ctx->launches++;
        }
        CK(cudaEventRecord(h.ev_out[g], ctx->stream));
        CK(cudaStreamWaitEvent(h.s_out, h.ev_out[g], 0));
        for (int j : ch) {
            int64_t r0, cnt; span(j, r0, cnt);
            CK(cudaMemcpyAsync(qout + r0 * U, ctx->stage + r0 * U, (size_t)cnt * U * sizeof(double), cudaMemcpyDeviceToHost, h.s_out));
        }
        return MSTGPU_OK;
    };
    for (int g = 0; g < G; g++) {
        int64_t r0, cnt; span(g, r0, cnt);
        CK(cudaMemcpyAsync(ctx->stage + r0 * U, qin + r0 * U, (size_t)cnt * U * sizeof(double), cudaMemcpyHostToDevice, h.s_in));
        CK(cudaEventRecord(h.ev_in[g], h.s_in));
        CK(cudaStreamWaitEvent(ctx->stream, h.ev_in[g], 0));
        k_rows_in<<<(unsigned)((cnt * U + 255) / 256), 256, 0, ctx->stream>>>(r0, cnt, U, ctx->stage, h.d_old2new, Qc);
        ctx->launches++;
        ctx->sub_group = g;
        r = launch_tiles_any<D>(ctx, dt, nullptr, Qc, Qn, 1, 0, ctx->stream);
        ctx->sub_group = -1;
        if (r) return r;
        if ((r = emit(g))) return r;
    }
    if (halo) {
        // every owned row is in place: ghost rows <- owners, then the tiles whose rings hold ghost cells
        if ((r = halo_exchange(ctx, Qc, ctx->stream))) return r;
        ctx->sub_group = G;
        r = launch_tiles_any<D>(ctx, dt, nullptr, Qc, Qn, 1, 1, ctx->stream);
        ctx->sub_group = -1;
        if (r) return r;
    }
    if ((r = emit(G))) return r;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(h.s_out));
    ctx->cur ^= 1;
    ctx->has_state = true;
    ctx->stepped = true;
    ctx->probes_valid = false;
    return MSTGPU_OK;
}

}  // namespace

template <int D>
static int step_implicit_impl(mstgpu_ctx* ctx, double dt, int nsteps, int iters) {
    constexpr int U = D + 2;
    const int nc = ctx->nc, nrow = ctx->n_owned;
    const bool dist = ctx->partitioned && !ctx->halo.empty();
    if (dist && !ctx->comm) { set_error(ctx, "partitioned context without a communicator: call mstgpu_comm_init"); return MSTGPU_ERR_STATE; }
    { int r0 = ensure_split_tables(ctx); if (r0) return r0; }  // the block assembly reads the per-face tables
    for (int s = 0; s < nsteps; s++) {
        double* Qc = ctx->Q[ctx->cur];
        double* Qn = ctx->Q[ctx->cur ^ 1];
        if (dist) {
            int r = halo_exchange(ctx, Qc, ctx->stream);
            if (r) return r;
        }
        // b = -R(Q): the explicit path's fluxes and gather in residual-vector mode
        if (ctx->use_tiles) {
            int r = launch_tiles_any<D>(ctx, dt, nullptr, Qc, ctx->imp_b, 2, 2, ctx->stream);
            if (r) return r;
        } else {
            int r = ensure_stage_buffers(ctx);
            if (r) return r;
            launch_gradient_stage<D>(ctx, Qc);
            const int nf = ctx->nf;
            {
                KTimer t(ctx, "flux");
                if (ctx->cfg.order == 2)
                    k_flux<D, 2><<<(nf + 127) / 128, 128, 0, ctx->stream>>>(nf, ctx->dcfg, Qc, ctx->G, ctx->fc0, ctx->fc1, ctx->meta, ctx->Sd,
                                                                             ctx->dx0, ctx->dx1, ctx->Phi, ctx->Gp, ctx->eta);
                else
                    k_flux<D, 1><<<(nf + 127) / 128, 128, 0, ctx->stream>>>(nf, ctx->dcfg, Qc, ctx->G, ctx->fc0, ctx->fc1, ctx->meta, ctx->Sd,
                                                                             ctx->dx0, ctx->dx1, ctx->Phi, ctx->Gp, ctx->eta);
            }
            KTimer t(ctx, "update");
            k_update<D><<<(nrow + 255) / 256, 256, 0, ctx->stream>>>(nc, nrow, ctx->nslot, 2, dt, nullptr, Qc, ctx->Phi, ctx->cf, ctx->vol,
                                                                     ctx->imp_b, ctx->resid, ctx->nanflag);
            ctx->probes_valid = true;
        }
        {
            KTimer t(ctx, "assemble");
            k_assemble_implicit<D><<<(nrow + 127) / 128, 128, 0, ctx->stream>>>(nrow, nc, ctx->nslot, ctx->dcfg.gamma, dt, Qc, ctx->cf, ctx->fc0,
                                                                                ctx->fc1, ctx->Sd, ctx->vol, ctx->imp_dpos, ctx->imp_pos, ctx->imp_val);
        }
        CK(cudaMemsetAsync(ctx->imp_x, 0, (size_t)nc * U * sizeof(double), ctx->stream));  // dQ starts from 0
        {
            KTimer t(ctx, "lusgs");
            const int64_t l0 = mstgpu_lusgs_launch_count(ctx->imp_solver);
            int r = MSTGPU_OK;
            lusgs_hint_zero_start(ctx->imp_solver);  // dQ starts from 0 (memset above)
            if (!dist) {
                r = lusgs_solve_async(ctx->imp_solver, ctx->stream, ctx->imp_val, ctx->imp_b, ctx->imp_x, iters, true);
            } else {
                // one sweep pair per call; the neighbours' dQ of this sweep become the next one's lagged ghost values
                for (int it = 0; it < iters && r == MSTGPU_OK; it++) {
                    r = lusgs_solve_async(ctx->imp_solver, ctx->stream, ctx->imp_val, ctx->imp_b, ctx->imp_x, 1, it == 0);
                    if (r == MSTGPU_OK && it + 1 < iters) {
                        r = halo_exchange(ctx, ctx->imp_x, ctx->stream);
                        if (r) return r;
                    }
                }
            }
            if (r) { set_error(ctx, std::string("lusgs: ") + mstgpu_lusgs_last_error()); return r; }
            ctx->launches += mstgpu_lusgs_launch_count(ctx->imp_solver) - l0 - 1;
        }
        CK(cudaMemsetAsync(ctx->resid, 0, 8 * sizeof(unsigned long long), ctx->stream));
        {
            KTimer t(ctx, "increment");
            k_add_increment<U><<<(nrow + 255) / 256, 256, 0, ctx->stream>>>(nrow, Qc, ctx->imp_x, Qn, ctx->resid, ctx->nanflag);
        }
        ctx->cur ^= 1;
    }
    CK(cudaGetLastError());
    ctx->stepped = nsteps > 0 || ctx->stepped;
    if (nsteps > 0 && ctx->use_tiles) ctx->probes_valid = false;
    return MSTGPU_OK;
}

// cells per tile: the largest tile that still lets two CTAs share an SM (228 KB of shared
// memory); the limiter extension adds a [U][own + ring 1] table, so its tiles are smaller
// staged packet stream: tets at second order without extensions, 128- or 256-thread CTAs
static bool tile_staged_for(const mstgpu_config& cfg, int D, int nslot) {
    if (!(D == 3 && nslot == 4 && cfg.order == 2 && !cfg.limiter && !cfg.viscous)) return false;
    if (cfg.block_threads != 0 && cfg.block_threads != 128 && cfg.block_threads != 256) return false;
    if (cfg.tile_flags & MSTGPU_TILE_DIRECT) return false;
    if (cfg.tile_flags & MSTGPU_TILE_STAGED) return true;
    static const char* env = getenv("MSTGPU_TILE_STAGED");
    return env ? atoi(env) != 0 : false;
}

// tets at second order without extensions (the scheme the metric is quoted on): 4 CTAs of 128 threads per SM on
// tiles of 240 cells instead of 2 x 256 threads on 512 -- the same 16 warps and registers, but four CTAs in
// different phases overlap the memory-bound phase 0 of one with the FP64-bound phase 2 of the others
// (measured at 50.2 M tets: 5.72-5.78 ms against 5.84-5.87 ms per step, profiles/r2_ab_variants.json)
static bool small_ctas_default(const mstgpu_config& cfg, int D, int nslot) {
    return D == 3 && nslot == 4 && cfg.order == 2 && !cfg.limiter && !cfg.viscous && cfg.block_threads == 0 && cfg.tile_cells == 0;
}

static int tile_threads_for(const mstgpu_config& cfg, int T, int D, int nslot) {
    if (small_ctas_default(cfg, D, nslot)) return 128;
    int NT = cfg.block_threads == 128 ? 128 : (cfg.block_threads == 256 ? 256 : (T <= 96 ? 128 : 256));
    if ((cfg.block_threads == 320 || cfg.block_threads == 384) && D == 3 && nslot == 4 && cfg.order == 2 && !cfg.limiter && !cfg.viscous)
        NT = cfg.block_threads;
    return NT;
}

// Variable tile sizes (tiles.h): flux faces per tile <= the returned count, 0 = fixed tile_cells.  Default for the
// 4 x 128-thread default on tets at second order: 512 faces = 4 trips of every warp through phase 2; fixed tiles of
// 240 cells carry ~595 faces on the 203^3 box: 5 trips, the last one with a handful of live warps (measured
// 5.83 -> 5.55 ms per step, profiles/r2_ab_variants.json).  MSTGPU_TILE_FIT overrides (experiments).
static int tile_fit_faces(const mstgpu_config& cfg, int D, int nslot) {
    if (const char* v = getenv("MSTGPU_TILE_FIT")) return std::max(0, atoi(v));
    if (cfg.tile_fit) return std::max(0, cfg.tile_fit);
    if (small_ctas_default(cfg, D, nslot)) return 512;
    // 2-D first order on triangles (256-thread CTAs, tiles of <= 512 cells carry ~820 faces: 4 trips, the last one
    // a fifth full): 768 faces = 3 full trips, 70.2 -> 64.8 us per step on 998 046 triangles (AUSM+)
    if (D == 2 && nslot == 3 && cfg.order == 1 && cfg.tile_cells == 0 && cfg.block_threads == 0) return 768;
    // limiter / viscous instantiations: measured neutral or slower (their per-cell phases grow with the rings of
    // smaller tiles), fixed tile_cells
    return 0;
}

static int default_tile_cells(const mstgpu_config& cfg, bool staged = false, int D = 0, int nslot = 0) {
    if (staged) return cfg.block_threads == 128 ? 208 : 416;  // + 80 B of landing slots per thread
    if (small_ctas_default(cfg, D, nslot)) return 256;  // cap; the tiles are sized by their flux faces (tile_fit_faces)
    if (cfg.order != 2) return 512;
    if (cfg.viscous != 0) return cfg.limiter != 0 ? 192 : 256;  // + [(D+1) D][own + ring 1] primitive gradients
    return cfg.limiter != 0 ? 384 : 512;
}

// how scattered are the ring rows of the tiles in memory: distinct 128-byte lines, runs of consecutive ring ids, and
// the 128-byte lines the gather of phase 0 requests (one thread per ring entry in list order, U loads of 8 bytes per
// thread: per load instruction the distinct lines among the 32 addresses of a warp = its L1 wavefronts)
static void tile_ring_runs(const TilePack& tp, int U, int64_t& rows_out, int64_t& lines_out, int64_t& runs_out, int64_t* gsec_out = nullptr) {
    const int64_t rowb = 8 * U;
    int64_t lines = 0, runs = 0, rows = 0, gsec = 0;
#pragma omp parallel for schedule(static) reduction(+ : lines, runs, rows, gsec)
    for (int t = 0; t < tp.ntiles; t++) {
        const TileDesc& d = tp.desc[t];
        const int n = d.n_r1 + d.n_r2;
        const int32_t* lst = tp.ring.data() + d.ring_off;
        for (int g = 0; g < n; g += 32)
            for (int k = 0; k < U; k++) {
                int64_t sec[32];
                const int m = std::min(32, n - g);
                for (int i = 0; i < m; i++) sec[i] = ((int64_t)lst[g + i] * rowb + 8 * k) / 128;
                std::sort(sec, sec + m);
                gsec += std::unique(sec, sec + m) - sec;
            }
        std::vector<int32_t> r(lst, lst + n);
        std::sort(r.begin(), r.end());
        int64_t last_line = -1;
        for (size_t i = 0; i < r.size(); i++) {
            if (i == 0 || r[i] != r[i - 1] + 1) runs++;
            const int64_t l0 = (int64_t)r[i] * rowb / 128, l1 = ((int64_t)r[i] * rowb + rowb - 1) / 128;
            for (int64_t l = std::max(l0, last_line + 1); l <= l1; l++) lines++;
            last_line = l1;
        }
        rows += (int64_t)r.size();
    }
    rows_out = rows; lines_out = lines; runs_out = runs;
    if (gsec_out) *gsec_out = gsec;
}

// The lattice of the Hilbert curve against the mesh.  The curve subdivides a cube; where its octree boxes fall on the
// cells decides the shape of the tiles (contiguous curve ranges) and how the ring cells of a tile are spread over memory.
// On lattice-like meshes this is not a small effect: on the 203^3 Kuhn box the runs of consecutive ring ids per tile
// vary between 88 and 136 with the extent of the cube (two partitions of the same mesh, cut from the same curve, ran
// 8 % apart for that reason alone: profiles/r2_scaling.md), and the flux faces per cell by 4 %.  The extent is therefore
// chosen by a search: ten cubes between 1x and 1.9x the bounding box (2x is the same lattice one level up), each
// evaluated by building the actual tiles on a sample -- a contiguous curve range of ~400 k cells from the middle of the
// mesh, cut out with its ghost layers like a partition -- and scored with a two-factor model of the kernel's time
// (flux faces per cell x ring lines per cell, below).  Unstructured meshes without a lattice score alike for every cube;
// the default cube then stays.  MSTGPU_CURVE_SEARCH=0 switches it off.
static CurveFrame choose_curve_frame(const mstgpu_mesh& m, const mstgpu_config& cfg) {
    CurveFrame base = bbox_frame(m, m.ncells);
    if (const char* v = getenv("MSTGPU_CURVE_SEARCH")) if (atoi(v) == 0) return base;
    if (getenv("MSTGPU_CURVE_SCALE")) return base;  // the experiment knob of plan.cpp fixes the cube itself
    if (cfg.renumber != 2 || cfg.kernel == 0 || !(base.ext > 0.0) || m.ncells < 4096) return base;
    const int nc = m.ncells, M = 400000;
    const mstgpu_mesh* sm = &m;
    Partition P;
    int n_own = -1;
    if (nc > M + M / 2) {
        std::vector<int32_t> ord;
        curve_order(m, 2, nc, ord, &base);
        std::vector<int32_t> cp(nc, 1);
        for (int i = nc / 2 - M / 2; i < nc / 2 + M / 2; i++) cp[ord[i]] = 0;
        if (!build_partition(m, cfg, 2, 0, cp.data(), P, &base).empty()) return base;
        sm = &P.mesh;
        n_own = P.n_owned;
    }
    mstgpu_config c2 = cfg;
    if (n_own >= 0) c2.qf_copy_from = 0x7fffffff;
    const bool verbose = getenv("MSTGPU_VERBOSE") != nullptr;
    double cost0 = 0.0;
    auto evaluate = [&](const CurveFrame& fr, const char* what, double arg) -> double {
        Plan p;
        p.frame = fr;
        if (!build_plan(*sm, c2, p, n_own).empty()) return -1.0;
        TilePack tp;
        const bool staged = tile_staged_for(cfg, p.D, p.nslot);
        const int T = cfg.tile_cells > 0 ? cfg.tile_cells : default_tile_cells(cfg, staged, p.D, p.nslot);
        int ext = tile_ext(cfg.order, cfg.limiter, cfg.viscous);
        if (staged) ext = tile_ext_staged(ext, tile_threads_for(cfg, T, p.D, p.nslot));
        const int nu = n_own >= 0 ? n_own : p.nc;
        if (!build_tiles(p, nu, T, cfg.order, tp, ext, tile_fit_faces(cfg, p.D, p.nslot)).empty()) return -1.0;
        int64_t rows, lines, runs;
        tile_ring_runs(tp, p.U, rows, lines, runs);
        const double F = (double)tp.sum_FB / nu, Ln = (double)lines / nu;
        // Calibrated on the B200 (12 cubes on the 101^3 and 128^3 Kuhn boxes, profiles/r2_curve_calibration.md): the
        // kernel's time follows (flux faces per cell) x (0.15 + distinct 128-byte lines of ring rows per cell) to 4 %.
        const double cost = F * (0.15 + Ln);
        if (cost0 == 0.0) cost0 = cost;
        if (verbose)
            fprintf(stderr, "[mstgpu] curve cube %s %.4g: %.3f flux faces, %.3f ring rows, %.3f ring lines, %.3f runs per cell -> cost %.4f\n", what, arg, F,
                    (double)rows / nu, Ln, (double)runs / nu, cost / cost0);
        return cost;
    };
    CurveFrame best_fr = base;
    double best_cost = evaluate(base, "x", 1.0);
    if (best_cost < 0.0) return base;
    // (a) ten extents of the bounding cube
    for (int k = 1; k < 10; k++) {
        CurveFrame fr = base;
        fr.ext = base.ext * (1.0 + 0.1 * k);
        const double c = evaluate(fr, "x", 1.0 + 0.1 * k);
        if (c > 0.0 && c < best_cost) { best_cost = c; best_fr = fr; }
    }
    // (b) cubes whose finest boxes ARE the mesh's own lattice, if it has one: a mesh made of k cells per lattice cell (6
    // Kuhn tets per hexahedron, 24 tets about a hexahedron's centroid, 2 triangles per quadrilateral ...) has the spacing
    // (V k / N)^(1/D); the octree is anchored at the domain's lower corner (the smallest face-centre coordinates: boundary
    // faces lie on it) and sized to a power of two of that spacing.  Boxes then hold whole lattice cells everywhere -- no
    // drift of the boxes against the mesh, which an extent of the bounding cube only achieves by accident (128^3).
    {
        const int D = m.dim;
        double V = 0.0;
        for (int c = 0; c < nc; c++) if (m.vol[c] == m.vol[c]) V += std::fabs(m.vol[c]);
        double flo[3] = {1e300, 1e300, 1e300}, fhi[3] = {-1e300, -1e300, -1e300};
        for (int f = 0; f < m.nfaces; f++)
            for (int d = 0; d < D; d++) {
                const double x = m.fc[(size_t)f * D + d];
                if (x == x) { flo[d] = std::min(flo[d], x); fhi[d] = std::max(fhi[d], x); }
            }
        double span = 0.0;
        for (int d = 0; d < D; d++) span = std::max(span, fhi[d] - flo[d]);
        if (V > 0.0 && span > 0.0)
            for (int k : {1, 2, 4, 5, 6, 8, 12, 24}) {
                const double h = std::pow(V * k / nc, 1.0 / D);
                double ext = h;
                while (ext < span * (1.0 + 1e-9)) ext *= 2.0;
                CurveFrame fr;
                fr.set = true;
                for (int d = 0; d < D; d++) fr.lo[d] = flo[d];
                fr.ext = ext;
                const double c = evaluate(fr, "lattice k =", (double)k);
                if (c > 0.0 && c < best_cost) { best_cost = c; best_fr = fr; }
            }
    }
    // the model is good to ~4 %: leave the bounding cube only for a predicted gain beyond that
    if (best_cost > 0.96 * cost0) return base;
    return best_fr;
}


extern "C" {

int mstgpu_tile_stats(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int64_t* out) { return mstgpu_tile_stats_owned(mesh, cfg, -1, out); }

int mstgpu_tile_stats_owned(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int32_t n_owned, int64_t* out) {
    if (!mesh || !cfg || !out) { set_error(nullptr, "null argument"); return MSTGPU_ERR_ARG; }
    if (n_owned > mesh->ncells) { set_error(nullptr, "n_owned > ncells"); return MSTGPU_ERR_ARG; }
    Plan p;
    if (n_owned < 0) p.frame = choose_curve_frame(*mesh, *cfg);
    std::string perr = build_plan(*mesh, *cfg, p, n_owned);
    if (!perr.empty()) { set_error(nullptr, perr); return MSTGPU_ERR_ARG; }
    TilePack tp;
    const bool staged = tile_staged_for(*cfg, p.D, p.nslot);
    int T = cfg->tile_cells > 0 ? cfg->tile_cells : default_tile_cells(*cfg, staged, p.D, p.nslot);
    const int NTs = tile_threads_for(*cfg, T, p.D, p.nslot);
    int ext = tile_ext(cfg->order, cfg->limiter, cfg->viscous);
    if (staged) ext = tile_ext_staged(ext, NTs);
    perr = build_tiles(p, n_owned >= 0 ? n_owned : p.nc, T, cfg->order, tp, ext, tile_fit_faces(*cfg, p.D, p.nslot));
    if (!perr.empty()) { set_error(nullptr, perr); return MSTGPU_ERR_ARG; }
    for (int i = 0; i < 16; i++) out[i] = 0;
    out[0] = tp.ntiles; out[1] = (int64_t)tp.max_smem;
    out[3] = tp.sum_r1; out[4] = tp.sum_r2; out[5] = tp.sum_FB; out[6] = tp.sum_FA; out[7] = (int64_t)tp.packets.size();
    double sum = 0;
    for (const TileDesc& d : tp.desc) {
        const size_t b = tile_layout(p.D, cfg->order, p.nslot, d.n_own, d.n_r1, d.n_r2, d.nFB, ext).total;
        sum += (double)b;
        out[8 + (b <= 56 * 1024 ? 0 : b <= 75 * 1024 ? 1 : b <= 113 * 1024 ? 2 : 3)]++;
        // loop trips of a CTA of NT threads: flux faces (phase 2), owned cells (phase 3), ring rows (phase 0)
        const int NT = NTs;
        out[12] += (d.nFB + NT - 1) / NT; out[13] += (d.n_own + NT - 1) / NT; out[14] += (d.n_r1 + d.n_r2 + NT - 1) / NT;
        out[15] = NT;
    }
    out[2] = (int64_t)(sum / tp.ntiles);
    return MSTGPU_OK;
}

int mstgpu_tile_locality(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int32_t n_owned, int64_t* out4) {  // out4: 6 entries
    if (!mesh || !cfg || !out4) { set_error(nullptr, "null argument"); return MSTGPU_ERR_ARG; }
    if (n_owned > mesh->ncells) { set_error(nullptr, "n_owned > ncells"); return MSTGPU_ERR_ARG; }
    Plan p;
    if (n_owned < 0) p.frame = choose_curve_frame(*mesh, *cfg);
    std::string perr = build_plan(*mesh, *cfg, p, n_owned);
    if (!perr.empty()) { set_error(nullptr, perr); return MSTGPU_ERR_ARG; }
    TilePack tp;
    int T = cfg->tile_cells > 0 ? cfg->tile_cells : default_tile_cells(*cfg, false, p.D, p.nslot);
    perr = build_tiles(p, n_owned >= 0 ? n_owned : p.nc, T, cfg->order, tp, tile_ext(cfg->order, cfg->limiter, cfg->viscous), tile_fit_faces(*cfg, p.D, p.nslot));
    if (!perr.empty()) { set_error(nullptr, perr); return MSTGPU_ERR_ARG; }
    int64_t lines = 0, runs = 0, rows = 0, gsec = 0;
    tile_ring_runs(tp, p.U, rows, lines, runs, &gsec);
    out4[0] = tp.ntiles; out4[1] = rows; out4[2] = lines; out4[3] = runs; out4[4] = gsec; out4[5] = tp.sum_FB;
    return MSTGPU_OK;
}

const char* mstgpu_version(void) { return "mstgpu 0.1 (sm_100a)"; }

const char* mstgpu_last_error(mstgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void mstgpu_default_config(mstgpu_config* cfg, int32_t dim) {
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->order = 2;
    cfg->flux = MSTGPU_FLUX_ROE;
    cfg->viscous = 0;
    cfg->qf_copy_from = -1;
    cfg->renumber = 2;
    cfg->device = -1;
    cfg->kernel = 1;
    cfg->tile_cells = 0;
    cfg->block_threads = 0;
    cfg->tile_flags = 0;
    cfg->gamma = 1.4;
    cfg->delta = 0.125;
    cfg->eor = 1e-10;
    cfg->mu = 1.7894e-05;
    cfg->kappa = 0.0242;
    cfg->cv = 715.8;
    cfg->gradient = MSTGPU_GRAD_GREEN_GAUSS;  // the reference's scheme: Green-Gauss, no limiter
    cfg->limiter = MSTGPU_LIMITER_NONE;
    cfg->limiter_k = 5.0;
    cfg->tile_fit = 0;
    cfg->reserved_ = 0;
    // CONST.h:70-83: rho = 1, u = v = w = 0, E = rho * (T*CV), T = 1/286.32
    cfg->inletQ[0] = 1.0;
    cfg->inletQ[dim + 1] = 1.0 * ((1 / 286.32) * 715.8 + 0.0);
}

static int create_impl(mstgpu_ctx** out, const mstgpu_mesh* mesh, const mstgpu_config* cfg, const Partition* part) {
    mstgpu_ctx* ctx = nullptr;
    if (!out || !mesh || !cfg) { set_error(nullptr, "null argument"); return MSTGPU_ERR_ARG; }
    *out = nullptr;
    if (cfg->order != 1 && cfg->order != 2) { set_error(nullptr, "order must be 1 or 2"); return MSTGPU_ERR_ARG; }
    if (cfg->flux != MSTGPU_FLUX_ROE && cfg->flux != MSTGPU_FLUX_AUSM) { set_error(nullptr, "unknown flux"); return MSTGPU_ERR_ARG; }
    if (cfg->viscous != 0 && cfg->viscous != 1) { set_error(nullptr, "viscous must be 0 or 1"); return MSTGPU_ERR_ARG; }
    if (cfg->gradient != MSTGPU_GRAD_GREEN_GAUSS && cfg->gradient != MSTGPU_GRAD_LSQ) { set_error(nullptr, "unknown gradient"); return MSTGPU_ERR_ARG; }
    if (cfg->limiter < 0 || cfg->limiter > MSTGPU_LIMITER_VENKATAKRISHNAN) { set_error(nullptr, "unknown limiter"); return MSTGPU_ERR_ARG; }
    if (cfg->limiter == MSTGPU_LIMITER_VENKATAKRISHNAN && !(cfg->limiter_k >= 0.0)) { set_error(nullptr, "limiter_k must be >= 0"); return MSTGPU_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error(nullptr, std::string("no CUDA device: ") + cudaGetErrorString(e));
        return MSTGPU_ERR_CUDA;
    }
    ctx = new mstgpu_ctx;
    ctx->cfg = *cfg;
    // the fused tile kernel carries the laminar viscous term at second order (it shares the rings of the
    // reconstruction); first order + viscous runs the split kernels
    ctx->use_tiles = cfg->kernel != 0 && (cfg->viscous == 0 || cfg->order == 2);
    mstgpu_config pcfg = *cfg;
    if (part) pcfg.qf_copy_from = 0x7fffffff;  // already folded into the partition's eta table
    // curve lattice: a partition orders its cells on the global mesh's lattice (so its tiles are the tiles the single-GPU
    // run has), a whole mesh on the cube the search picks
    if (part) ctx->plan.frame = part->frame;
    else {
        std::string verr = validate_mesh(*mesh);
        if (!verr.empty()) { set_error(nullptr, verr); delete ctx; return MSTGPU_ERR_ARG; }
        ctx->plan.frame = choose_curve_frame(*mesh, pcfg);
    }
    std::string perr = build_plan(*mesh, pcfg, ctx->plan, part ? part->n_owned : -1);
    if (!perr.empty()) { set_error(nullptr, perr); delete ctx; return MSTGPU_ERR_ARG; }
    Plan& p = ctx->plan;
    ctx->D = p.D; ctx->U = p.U; ctx->nc = p.nc; ctx->nf = p.nf; ctx->nslot = p.nslot;
    ctx->n_owned = part ? part->n_owned : p.nc;
    ctx->partitioned = part != nullptr;
    int rc = [&]() -> int {
        if (cfg->device >= 0) { CK(cudaSetDevice(cfg->device)); ctx->device = cfg->device; }
        else CK(cudaGetDevice(&ctx->device));
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        {
            // comm stream at the highest priority: when the exchange completes, the tiles that
            // were waiting for it are dispatched ahead of the interior launch's remaining CTAs
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
        }
        CK(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        CK(cudaEventCreate(&ctx->ev0));
        CK(cudaEventCreate(&ctx->ev1));
        int r;
        // the split kernels' tables (~230 B per cell on tets) go up now only if the split kernels are the step
        // path; under the fused kernel they wait for their first user (stage probes, CFL step, implicit step)
        if (!ctx->use_tiles && (r = ensure_split_tables(ctx))) return r;
        if ((r = upload(ctx, &ctx->cell_new2old, p.cell_new2old))) return r;
        const size_t nq = (size_t)p.nc * p.U;
        // + 2 rows: the fused kernel's bulk copies move an even number of rows
        if ((r = dalloc(ctx, &ctx->Q[0], nq + 2 * p.U))) return r;
        if ((r = dalloc(ctx, &ctx->Q[1], nq + 2 * p.U))) return r;
        CK(cudaMemsetAsync(ctx->Q[0], 0, (nq + 2 * p.U) * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(ctx->Q[1], 0, (nq + 2 * p.U) * sizeof(double), ctx->stream));
        if (ctx->use_tiles) {
            TilePack tp;
            ctx->tile_staged = tile_staged_for(*cfg, p.D, p.nslot);
            int T = cfg->tile_cells > 0 ? cfg->tile_cells : default_tile_cells(*cfg, ctx->tile_staged, p.D, p.nslot);
            ctx->tile_NT = tile_threads_for(*cfg, T, p.D, p.nslot);
            ctx->tile_ext = tile_ext(cfg->order, cfg->limiter, cfg->viscous);
            if (ctx->tile_staged) ctx->tile_ext = tile_ext_staged(ctx->tile_ext, ctx->tile_NT);
            std::string terr = build_tiles(p, ctx->n_owned, T, cfg->order, tp, ctx->tile_ext, tile_fit_faces(*cfg, p.D, p.nslot));
            if (!terr.empty()) { set_error(ctx, terr); return MSTGPU_ERR_ARG; }
            if (tp.open_stencils > 0) {
                // the fused kernel derives the own-cell reconstruction weight from the closure of the cell
                set_error(ctx, std::to_string(tp.open_stencils) + " face sides belong to cells that are not closed (sum of outward area vectors != 0): "
                               "the fused kernel needs closed cells; use kernel = 0 (split kernels) for this mesh");
                return MSTGPU_ERR_ARG;
            }
            int dev_smem = 0;
            CK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
            if (tp.max_smem + 1024 > (size_t)dev_smem) { set_error(ctx, "tile needs more shared memory than the device has; lower tile_cells"); return MSTGPU_ERR_ARG; }
            ctx->ntiles = tp.ntiles; ctx->tile_T = T; ctx->tile_smem = tp.max_smem;
            CK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
            if (const char* v = getenv("MSTGPU_TILE_VAR")) ctx->tile_var = atoi(v);
            {
                // The dynamic shared memory of a launch is what its LARGEST tile needs, and
                // it decides how many CTAs share an SM.  A few outlier tiles (ragged blobs)
                // must not cost every tile a resident CTA: group the tiles by the CTAs/SM
                // their size allows (4, 3, 2, 1) and launch each group with its own size.
                const size_t lim[4] = {57000, 76500, 115000, (size_t)dev_smem};
                std::vector<int> cls(tp.ntiles);
                size_t cmax[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int ccount[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int t = 0; t < tp.ntiles; t++) {
                    const TileDesc& d = tp.desc[t];
                    const size_t b = tile_layout(p.D, cfg->order, p.nslot, d.n_own, d.n_r1, d.n_r2, d.nFB, ctx->tile_ext).total;
                    int c = 0;
                    while (c < 3 && b > lim[c]) c++;
                    // tiles whose rings reach into the ghost cells wait for the halo exchange
                    bool halo = false;
                    for (int i = 0; i < d.n_r1 + d.n_r2 && !halo; i++) halo = tp.ring[d.ring_off + i] >= ctx->n_owned;
                    c += halo ? 4 : 0;
                    cls[t] = c; ccount[c]++; cmax[c] = std::max(cmax[c], b);
                }
                // merge a class that is too small to fill the machine into the next larger one
                for (int g = 0; g < 8; g += 4)
                    for (int c = g; c < g + 3; c++)
                        if (ccount[c] > 0 && ccount[c] < 64 && ccount[c + 1] > 0) {  // only a handful: not worth a launch
                            for (int t = 0; t < tp.ntiles; t++) if (cls[t] == c) cls[t] = c + 1;
                            ccount[c + 1] += ccount[c]; cmax[c + 1] = std::max(cmax[c + 1], cmax[c]); ccount[c] = 0;
                        }
                std::vector<TileDesc> sorted;
                sorted.reserve(tp.ntiles);
                for (int c = 0; c < 8; c++) {
                    if (!ccount[c]) continue;
                    ctx->tile_classes.push_back({(int)sorted.size(), ccount[c], cmax[c], c >= 4});
                    for (int t = 0; t < tp.ntiles; t++) if (cls[t] == c) sorted.push_back(tp.desc[t]);
                }
                tp.desc.swap(sorted);
                // ring ids at a fixed stride per tile, in the order of the sorted descriptors: the kernel loads them
                // together with the descriptor (step_tiles.cuh, phase 0)
                int stride = 4;
                for (const TileDesc& d : tp.desc) stride = std::max(stride, (d.n_r1 + d.n_r2 + 3) & ~3);
                std::vector<int32_t> ring((size_t)tp.ntiles * stride, 0);
#pragma omp parallel for schedule(static)
                for (int t = 0; t < tp.ntiles; t++) {
                    TileDesc& d = tp.desc[t];
                    std::copy(tp.ring.begin() + d.ring_off, tp.ring.begin() + d.ring_off + d.n_r1 + d.n_r2, ring.begin() + (size_t)t * stride);
                    d.ring_off = (int64_t)t * stride;
                }
                tp.ring.swap(ring);
                ctx->ring_stride = stride;
                // for the streamed step: the host row that completes a tile's input (largest reference-order row among
                // its owned cells and the ring cells this context owns; ghost rows arrive with the halo exchange)
                ctx->tile_ready_row.assign(tp.ntiles, 0);
                ctx->tile_cb_h.resize(tp.ntiles);
                ctx->tile_nown_h.resize(tp.ntiles);
#pragma omp parallel for schedule(static)
                for (int t = 0; t < tp.ntiles; t++) {
                    const TileDesc& d = tp.desc[t];
                    int32_t mx = 0;
                    for (int c = d.cb; c < d.cb + d.n_own; c++) mx = std::max(mx, p.cell_new2old[c]);
                    for (int i = 0; i < d.n_r1 + d.n_r2; i++) {
                        const int32_t g = tp.ring[d.ring_off + i];
                        if (g < ctx->n_owned) mx = std::max(mx, p.cell_new2old[g]);
                    }
                    ctx->tile_ready_row[t] = mx;
                    ctx->tile_cb_h[t] = d.cb;
                    ctx->tile_nown_h[t] = d.n_own;
                }
                if (getenv("MSTGPU_VERBOSE"))
                    for (const auto& tc : ctx->tile_classes)
                        fprintf(stderr, "[mstgpu] tile class: %d tiles, %zu B smem, %s\n", tc.count, tc.smem, tc.halo ? "halo" : "interior");
            }
            TileDesc* ddesc; int32_t* dring; unsigned char* dpk;
            if ((r = upload(ctx, &ddesc, tp.desc))) return r;
            ctx->tile_allocs.push_back(ddesc);
            if ((r = upload(ctx, &dring, tp.ring))) return r;
            ctx->tile_allocs.push_back(dring);
            if ((r = upload(ctx, &dpk, tp.packets))) return r;
            ctx->tile_allocs.push_back(dpk);
            ctx->ta = TileArrays{ddesc, dring, dpk, ctx->ring_stride, nullptr};
            CK(cudaStreamSynchronize(ctx->stream));  // tp goes out of scope
        }
        // staging buffer of the state permutation; the stage probes (gradient, face flux) grow it on first use
        if ((r = ensure_stage(ctx, nq))) return r;
        if (part) {
            std::vector<int32_t> sidx;
            for (const Neighbor& nb : part->nbrs) {
                mstgpu_ctx::HaloNb h{nb.rank, (int)sidx.size(), (int)nb.send_local.size(), nb.recv_first, nb.recv_count};
                for (int32_t l : nb.send_local) sidx.push_back(p.cell_old2new[l]);
                ctx->halo.push_back(h);
            }
            ctx->send_total = (int)sidx.size();
            if ((r = upload(ctx, &ctx->send_idx, sidx))) return r;
            if ((r = dalloc(ctx, &ctx->sendbuf, (size_t)std::max(1, ctx->send_total) * p.U))) return r;
        }
        if ((r = dalloc(ctx, &ctx->dtmin, (size_t)1))) return r;
        if ((r = dalloc(ctx, &ctx->dt_dev, (size_t)2))) return r;
        {
            const unsigned long long inf = 0x7FF0000000000000ULL;
            CK(cudaMemcpyAsync(ctx->dtmin, &inf, sizeof(inf), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemsetAsync(ctx->dt_dev, 0, 2 * sizeof(double), ctx->stream));
        }
        if ((r = dalloc(ctx, &ctx->resid, (size_t)8))) return r;
        if ((r = dalloc(ctx, &ctx->nanflag, (size_t)1))) return r;
        CK(cudaMemsetAsync(ctx->resid, 0, 8 * sizeof(unsigned long long), ctx->stream));
        CK(cudaMemsetAsync(ctx->nanflag, 0, sizeof(int), ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return MSTGPU_OK;
    }();
    if (rc != MSTGPU_OK) {
        g_create_error = ctx->err;
        mstgpu_destroy(ctx);
        return rc;
    }
    // device config
    DevCfg& d = ctx->dcfg;
    d.gamma = cfg->gamma; d.gm1 = cfg->gamma - 1.0; d.delta = cfg->delta;
    d.delta2 = cfg->delta * cfg->delta; d.inv2delta = 1.0 / (2.0 * cfg->delta); d.eor = cfg->eor;
    d.astar_fac = 2.0 * (cfg->gamma - 1.0) / (cfg->gamma + 1.0);
    d.mu = cfg->mu; d.lambda = -0.666667 * cfg->mu; d.kappa = cfg->kappa; d.cv = cfg->cv; d.inv_cv = 1.0 / cfg->cv;
    for (int k = 0; k < 5; k++) d.inletQ[k] = cfg->inletQ[k];
    d.order = cfg->order; d.flux = cfg->flux; d.viscous = cfg->viscous; d.limiter = cfg->order == 2 ? cfg->limiter : 0;
    // the big host tables are no longer needed once they are on the device (ensure_split_tables drops them)
    *out = ctx;
    return MSTGPU_OK;
}

int mstgpu_create(mstgpu_ctx** out, const mstgpu_mesh* mesh, const mstgpu_config* cfg) {
    return create_impl(out, mesh, cfg, nullptr);
}

void mstgpu_destroy(mstgpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void* q : ctx->peer_opened) cudaIpcCloseMemHandle(q);
    for (void* q : {(void*)ctx->peer_flags, (void*)ctx->push_dst, (void*)ctx->push_slot, (void*)ctx->push_ticket, (void*)ctx->peer_epochs})
        if (q) cudaFree(q);
    if (ctx->comm && g_nccl.h) g_nccl.CommDestroy(ctx->comm);
    if (ctx->send_idx) cudaFree(ctx->send_idx);
    if (ctx->sendbuf) cudaFree(ctx->sendbuf);
    void* ptrs[] = {ctx->Q[0], ctx->Q[1], ctx->G, ctx->Gp, ctx->Phi, ctx->stage, ctx->Sd, ctx->dx0, ctx->dx1, ctx->eta,
                    ctx->vol, ctx->fc0, ctx->fc1, ctx->cf, ctx->cell_new2old, ctx->face_new2old, ctx->meta,
                    ctx->resid, ctx->nanflag, ctx->lsq, ctx->eps2, ctx->dtmin, ctx->dt_dev, ctx->imp_dpos, ctx->imp_pos,
                    ctx->imp_val, ctx->imp_b, ctx->imp_x, ctx->out_nf_ptr, ctx->out_nf_idx, ctx->out_c0, ctx->out_c1,
                    ctx->out_eta, ctx->out_w, ctx->out_fields, ctx->out_Q, ctx->out_send_idx, ctx->out_sendbuf};
    if (ctx->imp_solver) mstgpu_lusgs_destroy(ctx->imp_solver);
    for (void* q : ptrs)
        if (q) cudaFree(q);
    for (void* q : ctx->tile_allocs)
        if (q) cudaFree(q);
    for (auto& pnd : ctx->pending) { cudaEventDestroy(pnd.second.first); cudaEventDestroy(pnd.second.second); }
    for (auto e : ctx->evpool) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    host_stream_free(ctx);
    for (auto ge : ctx->step_graph) if (ge) cudaGraphExecDestroy(ge);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int mstgpu_set_state(mstgpu_ctx* ctx, const double* q, int64_t ncells) {
    if (!ctx || !q) return MSTGPU_ERR_ARG;
    if (ncells != ctx->n_owned) { set_error(ctx, "set_state: ncells mismatch"); return MSTGPU_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    const size_t tot = (size_t)ctx->n_owned * ctx->U;
    CK(cudaMemcpyAsync(ctx->stage, q, tot * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_permute_in<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ctx->n_owned, ctx->U, ctx->stage,
                                                                         ctx->cell_new2old, ctx->Q[ctx->cur]);
    ctx->launches++;
    CK(cudaGetLastError());
    // prev state == current until the first step (Time.cpp:16-18 fills both)
    CK(cudaMemcpyAsync(ctx->Q[ctx->cur ^ 1], ctx->Q[ctx->cur], (size_t)ctx->nc * ctx->U * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->nanflag, 0, sizeof(int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->has_state = true;
    return MSTGPU_OK;
}

int mstgpu_get_state(mstgpu_ctx* ctx, double* q) {
    if (!ctx || !q) return MSTGPU_ERR_ARG;
    if (!ctx->has_state) { set_error(ctx, "get_state before set_state"); return MSTGPU_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    return fetch_permuted(ctx, ctx->Q[ctx->cur], ctx->cell_new2old, ctx->n_owned, ctx->U, q);
}

int mstgpu_get_prev_state(mstgpu_ctx* ctx, double* q) {
    if (!ctx || !q) return MSTGPU_ERR_ARG;
    if (!ctx->has_state) { set_error(ctx, "get_prev_state before set_state"); return MSTGPU_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    return fetch_permuted(ctx, ctx->Q[ctx->cur ^ 1], ctx->cell_new2old, ctx->n_owned, ctx->U, q);
}

int mstgpu_step(mstgpu_ctx* ctx, double dt, int32_t nsteps) {
    if (!ctx) return MSTGPU_ERR_ARG;
    if (!ctx->has_state) { set_error(ctx, "step before set_state"); return MSTGPU_ERR_STATE; }
    if (nsteps < 0) { set_error(ctx, "nsteps < 0"); return MSTGPU_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    int rc = (ctx->D == 2) ? step_impl<2>(ctx, dt, nsteps) : step_impl<3>(ctx, dt, nsteps);
    if (ctx->ktiming) drain_timers(ctx);
    return rc;
}

int mstgpu_step_host(mstgpu_ctx* ctx, const double* q_in, double* q_out, double dt, int32_t nchunks) {
    if (!ctx || !q_in || !q_out) return MSTGPU_ERR_ARG;
    if (!ctx->use_tiles) { set_error(ctx, "step_host needs the fused kernel (kernel = 1)"); return MSTGPU_ERR_STATE; }
    if (ctx->tile_var & 3) { set_error(ctx, "step_host: not with the persistent / prefetch-ahead launch variants"); return MSTGPU_ERR_STATE; }
    if (ctx->partitioned && !ctx->halo.empty() && !ctx->comm) { set_error(ctx, "partitioned context without a communicator: call mstgpu_comm_init"); return MSTGPU_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    if (nchunks <= 0) {
        nchunks = 64;
        if (const char* v = getenv("MSTGPU_HOST_CHUNKS")) nchunks = std::max(1, atoi(v));
    }
    int r = ensure_stage(ctx, (size_t)ctx->n_owned * ctx->U);
    if (r) return r;
    if ((r = host_stream_setup(ctx, nchunks))) return r;
    return (ctx->D == 2) ? step_host_impl<2>(ctx, q_in, q_out, dt) : step_host_impl<3>(ctx, q_in, q_out, dt);
}

int mstgpu_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return MSTGPU_ERR_ARG;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return MSTGPU_ERR_CUDA; }
    return MSTGPU_OK;
}

int mstgpu_host_unregister(void* p) {
    if (!p) return MSTGPU_OK;
    const cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return MSTGPU_ERR_CUDA; }
    return MSTGPU_OK;
}

int mstgpu_step_timed(mstgpu_ctx* ctx, double dt, int32_t nsteps, float* ms) {
    if (!ctx || !ms) return MSTGPU_ERR_ARG;
    if (!ctx->has_state) { set_error(ctx, "step before set_state"); return MSTGPU_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    // one-time work (graph capture + instantiation on first use) stays outside the timed bracket
    if (ctx->D == 2) step_graph_ready<2>(ctx, dt, nsteps, 0.0); else step_graph_ready<3>(ctx, dt, nsteps, 0.0);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc = (ctx->D == 2) ? step_impl<2>(ctx, dt, nsteps) : step_impl<3>(ctx, dt, nsteps);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    if (ctx->ktiming) drain_timers(ctx);
    return MSTGPU_OK;
}

int mstgpu_cfl_dt(mstgpu_ctx* ctx, double cfl, double* dt) {
    if (!ctx || !dt) return MSTGPU_ERR_ARG;
    if (!ctx->has_state) { set_error(ctx, "cfl_dt before set_state"); return MSTGPU_ERR_STATE; }
    if (!(cfl > 0.0)) { set_error(ctx, "cfl must be > 0"); return MSTGPU_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    int rc = (ctx->D == 2) ? cfl_on_device<2>(ctx, cfl, ctx->Q[ctx->cur]) : cfl_on_device<3>(ctx, cfl, ctx->Q[ctx->cur]);
    if (rc) return rc;
    double h[2];
    CK(cudaMemcpyAsync(h, ctx->dt_dev, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    // a query does not advance the clock of mstgpu_step_cfl
    CK(cudaMemsetAsync(ctx->dt_dev + 1, 0, sizeof(double), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *dt = h[0];
    return MSTGPU_OK;
}

int mstgpu_step_cfl(mstgpu_ctx* ctx, double cfl, int32_t nsteps, double* time_advanced) {
    return mstgpu_step_cfl_timed(ctx, cfl, nsteps, time_advanced, nullptr);
}

int mstgpu_step_cfl_timed(mstgpu_ctx* ctx, double cfl, int32_t nsteps, double* time_advanced, float* ms) {
    if (!ctx) return MSTGPU_ERR_ARG;
    if (!ctx->has_state) { set_error(ctx, "step before set_state"); return MSTGPU_ERR_STATE; }
    if (nsteps < 0) { set_error(ctx, "nsteps < 0"); return MSTGPU_ERR_ARG; }
    if (!(cfl > 0.0)) { set_error(ctx, "cfl must be > 0"); return MSTGPU_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->dt_dev + 1, 0, sizeof(double), ctx->stream));
    if (ms) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
    }
    int rc = (ctx->D == 2) ? step_impl<2>(ctx, 0.0, nsteps, cfl) : step_impl<3>(ctx, 0.0, nsteps, cfl);
    if (rc == MSTGPU_OK && ms) {
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    }
    if (ctx->ktiming) drain_timers(ctx);
    if (rc) return rc;
    if (time_advanced) {
        CK(cudaMemcpyAsync(time_advanced, ctx->dt_dev + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return MSTGPU_OK;
}

// ---- implicit step (extension; SURVEY.md 8a row L: the reference has the LU-SGS solver but no
// rhoSolver call site, so the operator is build-defined -- see oracle/rho_oracle.cpp implicitSystem) ----
int mstgpu_implicit_setup(mstgpu_ctx* ctx, int32_t colour_sweeps) {
    if (!ctx) return MSTGPU_ERR_ARG;
    if (ctx->imp_solver) return MSTGPU_OK;
    CK(cudaSetDevice(ctx->device));
    { int r0 = ensure_split_tables(ctx); if (r0) return r0; }
    // rows = the cells this context advances; on a partition the ghost cells (ids >= n_owned) are
    // columns only, their couplings lagged by one sweep (block Jacobi across partitions)
    const int nc = ctx->nc, nrow = ctx->n_owned, nslot = ctx->nslot, U = ctx->U;
    // the connectivity lives on the device (the host copies were dropped after upload)
    std::vector<int32_t> cf((size_t)nslot * nc), fc0(ctx->nf), fc1(ctx->nf);
    CK(cudaMemcpy(cf.data(), ctx->cf, cf.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fc0.data(), ctx->fc0, fc0.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fc1.data(), ctx->fc1, fc1.size() * 4, cudaMemcpyDeviceToHost));
    auto nb_of = [&](int c, int j) -> int {
        const int v = cf[(size_t)j * nc + c];
        if (v < 0) return -1;
        return (v & 1) ? fc0[v >> 1] : fc1[v >> 1];
    };
    std::vector<int32_t> rowptr((size_t)nrow + 1, 0);
    for (int c = 0; c < nrow; c++) {
        int cnt = 1;
        for (int j = 0; j < nslot; j++) cnt += nb_of(c, j) >= 0 ? 1 : 0;
        const int64_t next = (int64_t)rowptr[c] + cnt;
        if (next > 0x7fffffffLL) { set_error(ctx, "implicit system does not fit 32-bit offsets"); return MSTGPU_ERR_ARG; }
        rowptr[c + 1] = (int32_t)next;
    }
    const size_t nnz = (size_t)rowptr[nrow];
    std::vector<int32_t> col(nnz), dpos(nrow), pos((size_t)nslot * nc, -1);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nrow; c++) {
        int32_t* out = col.data() + rowptr[c];
        int n = 0;
        out[n++] = c;
        for (int j = 0; j < nslot; j++) { const int nb = nb_of(c, j); if (nb >= 0) out[n++] = nb; }
        std::sort(out, out + n);
        dpos[c] = rowptr[c] + (int32_t)(std::lower_bound(out, out + n, c) - out);
        for (int j = 0; j < nslot; j++) {
            const int nb = nb_of(c, j);
            if (nb >= 0) pos[(size_t)j * nc + c] = rowptr[c] + (int32_t)(std::lower_bound(out, out + n, nb) - out);
        }
    }
    std::vector<int32_t> order;
    if (colour_sweeps) {
        order.resize(nrow);
        int32_t ncol = 0;
        if (mstgpu_lusgs_color_order_partitioned(nrow, nc, rowptr.data(), col.data(), order.data(), &ncol) != MSTGPU_OK) {
            set_error(ctx, std::string("colour order: ") + mstgpu_lusgs_last_error());
            return MSTGPU_ERR_ARG;
        }
    }
    int rc = mstgpu_lusgs_create_partitioned(&ctx->imp_solver, nrow, nc, U, rowptr.data(), col.data(),
                                             colour_sweeps ? order.data() : nullptr, ctx->device);
    if (rc != MSTGPU_OK) { set_error(ctx, std::string("lusgs: ") + mstgpu_lusgs_last_error()); return rc; }
    const int64_t bytes_before = ctx->dev_bytes;
    ctx->dev_bytes += mstgpu_lusgs_device_bytes(ctx->imp_solver);
    ctx->imp_sweep = order;
    // imp_solver != nullptr means "fully set up" to mstgpu_step_implicit: on any failure from here on (out of
    // memory on the block array at tens of millions of rows is plausible) everything is undone
    rc = [&]() -> int {
        int r;
        if ((r = upload(ctx, &ctx->imp_dpos, dpos))) return r;
        if ((r = upload(ctx, &ctx->imp_pos, pos))) return r;
        if ((r = dalloc(ctx, &ctx->imp_val, nnz * U * U))) return r;
        if ((r = dalloc(ctx, &ctx->imp_b, (size_t)nc * U + 2 * U))) return r;  // + 2 rows: bulk stores of the fused kernel
        CK(cudaMemsetAsync(ctx->imp_b, 0, ((size_t)nc * U + 2 * U) * sizeof(double), ctx->stream));
        if ((r = dalloc(ctx, &ctx->imp_x, (size_t)nc * U))) return r;
        CK(cudaStreamSynchronize(ctx->stream));
        return MSTGPU_OK;
    }();
    if (rc != MSTGPU_OK) {
        cudaGetLastError();  // clear the sticky allocation error: the context itself stays usable
        mstgpu_lusgs_destroy(ctx->imp_solver);
        ctx->imp_solver = nullptr;
        if (ctx->imp_dpos) { cudaFree(ctx->imp_dpos); ctx->imp_dpos = nullptr; }
        if (ctx->imp_pos) { cudaFree(ctx->imp_pos); ctx->imp_pos = nullptr; }
        if (ctx->imp_val) { cudaFree(ctx->imp_val); ctx->imp_val = nullptr; }
        if (ctx->imp_b) { cudaFree(ctx->imp_b); ctx->imp_b = nullptr; }
        if (ctx->imp_x) { cudaFree(ctx->imp_x); ctx->imp_x = nullptr; }
        ctx->imp_sweep.clear();
        ctx->dev_bytes = bytes_before;
    }
    return rc;
}

int mstgpu_implicit_sweep_order(mstgpu_ctx* ctx, int32_t* order_ref_ids) {
    if (!ctx || !order_ref_ids) return MSTGPU_ERR_ARG;
    if (!ctx->imp_solver) { set_error(ctx, "implicit_sweep_order before implicit_setup"); return MSTGPU_ERR_STATE; }
    for (int i = 0; i < ctx->n_owned; i++) {  // a partition's ids are its own (mstgpu_partition_cell_ids maps them)
        const int dev = ctx->imp_sweep.empty() ? i : ctx->imp_sweep[i];
        order_ref_ids[i] = ctx->plan.cell_new2old[dev];
    }
    return MSTGPU_OK;
}

int mstgpu_step_implicit(mstgpu_ctx* ctx, double dt, int32_t nsteps, int32_t lusgs_iters, float* ms) {
    if (!ctx) return MSTGPU_ERR_ARG;
    if (!ctx->has_state) { set_error(ctx, "step before set_state"); return MSTGPU_ERR_STATE; }
    if (nsteps < 0 || lusgs_iters < 0 || !(dt > 0.0)) { set_error(ctx, "step_implicit: bad dt / nsteps / iterations"); return MSTGPU_ERR_ARG; }
    if (!ctx->imp_solver) {
        int r = mstgpu_implicit_setup(ctx, 1);
        if (r) return r;
    }
    CK(cudaSetDevice(ctx->device));
    if (ms) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
    }
    int rc = (ctx->D == 2) ? step_implicit_impl<2>(ctx, dt, nsteps, lusgs_iters) : step_implicit_impl<3>(ctx, dt, nsteps, lusgs_iters);
    if (rc == MSTGPU_OK && ms) {
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    }
    if (ctx->ktiming) drain_timers(ctx);
    return rc;
}

int mstgpu_sync(mstgpu_ctx* ctx) {
    if (!ctx) return MSTGPU_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return MSTGPU_OK;
}

int mstgpu_residual_linf(mstgpu_ctx* ctx, double* out) {
    if (!ctx || !out) return MSTGPU_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    unsigned long long bits[8];
    int nan = 0;
    if (ctx->comm) {
        // max of non-negative doubles == max of their bit patterns as unsigned integers
        NK(g_nccl.AllReduce(ctx->resid, ctx->resid, 8, ncclUint64, ncclMax, ctx->comm, ctx->stream));
        NK(g_nccl.AllReduce(ctx->nanflag, ctx->nanflag, 1, ncclInt32, ncclMax, ctx->comm, ctx->stream));
    }
    CK(cudaMemcpyAsync(bits, ctx->resid, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&nan, ctx->nanflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < ctx->U; k++) std::memcpy(&out[k], &bits[k], 8);
    if (nan & 4) { set_error(ctx, "peer halo: a neighbour's rows did not arrive within the wait bound (lost rank?)"); return MSTGPU_ERR_NCCL; }
    if (nan) { set_error(ctx, "NaN in the state"); return MSTGPU_ERR_NAN; }
    return MSTGPU_OK;
}

int mstgpu_debug_gradient(mstgpu_ctx* ctx, double* grad) {
    if (!ctx || !grad) return MSTGPU_ERR_ARG;
    if (ctx->cfg.order != 2 || !ctx->stepped) { set_error(ctx, "no gradient: order 1 or no step yet"); return MSTGPU_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    if (!ctx->probes_valid) {
        int r = (ctx->D == 2) ? recompute_stages<2>(ctx) : recompute_stages<3>(ctx);
        if (r) return r;
    }
    return fetch_permuted(ctx, ctx->G, ctx->cell_new2old, ctx->nc, ctx->U * ctx->D, grad);
}

int mstgpu_debug_face_flux(mstgpu_ctx* ctx, double* phi) {
    if (!ctx || !phi) return MSTGPU_ERR_ARG;
    if (!ctx->stepped) { set_error(ctx, "no flux: no step yet"); return MSTGPU_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    if (!ctx->probes_valid) {
        int r = (ctx->D == 2) ? recompute_stages<2>(ctx) : recompute_stages<3>(ctx);
        if (r) return r;
    }
    return fetch_permuted(ctx, ctx->Phi, ctx->face_new2old, ctx->nf, ctx->U, phi);
}

int64_t mstgpu_launch_count(mstgpu_ctx* ctx) { return ctx ? ctx->launches : -1; }

int mstgpu_set_tile_variant(mstgpu_ctx* ctx, int32_t variant) {
    if (!ctx) return MSTGPU_ERR_ARG;
    const int base = variant & ~(256 | 512 | 1024);  // + packet-stream experiments of the 128-thread default (step_tiles.cuh)
    if (variant < 0 || !(base == 8 || base == 16 || base == 32 || base == 64 || (base >= 0 && base <= 7 && base != 6))) { set_error(ctx, "tile variant must be 0-5, 7, 8, 16, 32 or 64 (+ 256 / 512 / 1024)"); return MSTGPU_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& ge : ctx->step_graph) if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }  // the graph holds the old kernels
    ctx->tile_var = variant;
    return MSTGPU_OK;
}

int mstgpu_enable_kernel_timing(mstgpu_ctx* ctx, int32_t on) {
    if (!ctx) return MSTGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    drain_timers(ctx);
    ctx->ktiming = on != 0;
    ctx->kstat.clear();
    return MSTGPU_OK;
}

int mstgpu_kernel_time(mstgpu_ctx* ctx, const char* name, double* ms, int64_t* launches) {
    if (!ctx || !name) return MSTGPU_ERR_ARG;
    auto it = ctx->kstat.find(name);
    if (it == ctx->kstat.end()) { if (ms) *ms = 0; if (launches) *launches = 0; return MSTGPU_OK; }
    if (ms) *ms = it->second.ms;
    if (launches) *launches = it->second.launches;
    return MSTGPU_OK;
}

int64_t mstgpu_device_bytes(mstgpu_ctx* ctx) { return ctx ? ctx->dev_bytes : -1; }

struct mstgpu_part {
    Partition p;
};

int mstgpu_partition_create(mstgpu_part** out, const mstgpu_mesh* g, const mstgpu_config* cfg, int32_t nparts,
                            int32_t rank, const int32_t* cell_part) {
    if (!out || !g || !cfg) { set_error(nullptr, "null argument"); return MSTGPU_ERR_ARG; }
    *out = nullptr;
    mstgpu_part* h = new mstgpu_part;
    {
        std::string verr = validate_mesh(*g);
        if (!verr.empty()) { set_error(nullptr, verr); delete h; return MSTGPU_ERR_ARG; }
    }
    const CurveFrame fr = choose_curve_frame(*g, *cfg);
    std::string e = build_partition(*g, *cfg, nparts, rank, cell_part, h->p, &fr);
    if (!e.empty()) { set_error(nullptr, e); delete h; return MSTGPU_ERR_ARG; }
    *out = h;
    return MSTGPU_OK;
}
void mstgpu_partition_destroy(mstgpu_part* part) { delete part; }
const mstgpu_mesh* mstgpu_partition_mesh(const mstgpu_part* part) { return part ? &part->p.mesh : nullptr; }
int mstgpu_partition_sizes(const mstgpu_part* part, int32_t* n_owned, int32_t* n_local, int32_t* n_neighbors) {
    if (!part) return MSTGPU_ERR_ARG;
    if (n_owned) *n_owned = part->p.n_owned;
    if (n_local) *n_local = part->p.n_local;
    if (n_neighbors) *n_neighbors = (int32_t)part->p.nbrs.size();
    return MSTGPU_OK;
}
const int32_t* mstgpu_partition_cell_ids(const mstgpu_part* part) { return part ? part->p.local2global.data() : nullptr; }
int mstgpu_partition_neighbor(const mstgpu_part* part, int32_t i, int32_t* rank, int32_t* send_count,
                              const int32_t** send_local, int32_t* recv_first, int32_t* recv_count) {
    if (!part || i < 0 || i >= (int32_t)part->p.nbrs.size()) return MSTGPU_ERR_ARG;
    const Neighbor& n = part->p.nbrs[i];
    if (rank) *rank = n.rank;
    if (send_count) *send_count = (int32_t)n.send_local.size();
    if (send_local) *send_local = n.send_local.data();
    if (recv_first) *recv_first = n.recv_first;
    if (recv_count) *recv_count = n.recv_count;
    return MSTGPU_OK;
}

int mstgpu_create_partitioned(mstgpu_ctx** out, const mstgpu_part* part, const mstgpu_config* cfg) {
    if (!part) { set_error(nullptr, "null partition"); return MSTGPU_ERR_ARG; }
    return create_impl(out, &part->p.mesh, cfg, &part->p);
}

int mstgpu_comm_unique_id(char* out128) {
    mstgpu_ctx* ctx = nullptr;
    if (!out128) return MSTGPU_ERR_ARG;
    if (!g_nccl.load()) { set_error(nullptr, g_nccl.err); return MSTGPU_ERR_NCCL; }
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, 128);
    return MSTGPU_OK;
}

int mstgpu_comm_init(mstgpu_ctx* ctx, int32_t nranks, int32_t rank, const char* id128) {
    if (!ctx || !id128) return MSTGPU_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!g_nccl.load()) { set_error(ctx, g_nccl.err); return MSTGPU_ERR_NCCL; }
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    NK(g_nccl.CommInitRank(&ctx->comm, nranks, id, rank));
    ctx->nranks = nranks; ctx->rank = rank;
    return MSTGPU_OK;
}

// ---- peer-memory halo: handle exchange ------------------------------------------------------------------
// Every rank exports one fixed-size blob (IPC handles of its two state buffers and of its flag array, plus
// where each neighbour's rows go in ITS local numbering); the host all-gathers the blobs by whatever means it
// has (torch.distributed, MPI, a file) and every rank opens its neighbours' buffers.
struct PeerBlob {
    cudaIpcMemHandle_t q[2], flags;
    int32_t rank, nnb;
    int32_t nb_rank[MSTGPU_MAX_NB], nb_recv_first[MSTGPU_MAX_NB], nb_recv_count[MSTGPU_MAX_NB];
};

int64_t mstgpu_peer_blob_bytes(void) { return (int64_t)sizeof(PeerBlob); }

int mstgpu_peer_export(mstgpu_ctx* ctx, int32_t rank, void* blob) {
    if (!ctx || !blob || rank < 0) return MSTGPU_ERR_ARG;
    if (!ctx->partitioned) { set_error(ctx, "peer_export: not a partitioned context"); return MSTGPU_ERR_STATE; }
    if ((int)ctx->halo.size() > MSTGPU_MAX_NB) { set_error(ctx, "peer halo supports up to 32 neighbours per rank"); return MSTGPU_ERR_ARG; }
    CK(cudaSetDevice(ctx->device));
    if (!ctx->peer_flags) {
        int r;
        if ((r = dalloc(ctx, &ctx->peer_flags, (size_t)MSTGPU_MAX_NB))) return r;
        if ((r = dalloc(ctx, &ctx->push_ticket, (size_t)1))) return r;
        if ((r = dalloc(ctx, &ctx->peer_epochs, (size_t)2))) return r;
        CK(cudaMemset(ctx->peer_epochs, 0, 2 * sizeof(unsigned long long)));
        CK(cudaMemset(ctx->peer_flags, 0, MSTGPU_MAX_NB * sizeof(unsigned long long)));
        CK(cudaMemset(ctx->push_ticket, 0, sizeof(unsigned int)));
    }
    PeerBlob b;
    std::memset(&b, 0, sizeof(b));
    CK(cudaIpcGetMemHandle(&b.q[0], ctx->Q[0]));
    CK(cudaIpcGetMemHandle(&b.q[1], ctx->Q[1]));
    CK(cudaIpcGetMemHandle(&b.flags, ctx->peer_flags));
    b.rank = rank; b.nnb = (int32_t)ctx->halo.size();
    for (size_t i = 0; i < ctx->halo.size(); i++) {
        b.nb_rank[i] = ctx->halo[i].rank; b.nb_recv_first[i] = ctx->halo[i].recv_first; b.nb_recv_count[i] = ctx->halo[i].recv_count;
    }
    std::memcpy(blob, &b, sizeof(b));
    return MSTGPU_OK;
}

int mstgpu_peer_connect(mstgpu_ctx* ctx, int32_t nranks, int32_t rank, const void* blobs) {
    if (!ctx || !blobs || nranks < 1 || rank < 0 || rank >= nranks) return MSTGPU_ERR_ARG;
    if (!ctx->partitioned || !ctx->peer_flags) { set_error(ctx, "peer_connect before peer_export"); return MSTGPU_ERR_STATE; }
    if (ctx->peer_ok) return MSTGPU_OK;
    CK(cudaSetDevice(ctx->device));
    const PeerBlob* all = static_cast<const PeerBlob*>(blobs);
    PeerTable pt;
    std::memset(&pt, 0, sizeof(pt));
    pt.nnb = (int)ctx->halo.size();
    std::vector<int32_t> dst((size_t)std::max(1, ctx->send_total));
    std::vector<uint8_t> slot((size_t)std::max(1, ctx->send_total));
    auto fail = [&](const std::string& m) {
        for (void* q : ctx->peer_opened) cudaIpcCloseMemHandle(q);
        ctx->peer_opened.clear();
        cudaGetLastError();
        set_error(ctx, m);
        return MSTGPU_ERR_CUDA;
    };
    for (size_t i = 0; i < ctx->halo.size(); i++) {
        const auto& h = ctx->halo[i];
        if (h.rank < 0 || h.rank >= nranks) return fail("peer_connect: neighbour rank out of range");
        const PeerBlob& nb = all[h.rank];
        if (nb.rank != h.rank) return fail("peer_connect: blob " + std::to_string(h.rank) + " was exported by rank " + std::to_string(nb.rank));
        int mine = -1;
        for (int j = 0; j < nb.nnb; j++) if (nb.nb_rank[j] == rank) mine = j;
        if (mine < 0 && (h.send_count || h.recv_count)) return fail("peer_connect: neighbour does not list this rank");
        if (mine >= 0 && nb.nb_recv_count[mine] != h.send_count) return fail("peer_connect: send / receive counts of a neighbour pair differ");
        void *q0 = nullptr, *q1 = nullptr, *fl = nullptr;
        cudaError_t e;
        if ((e = cudaIpcOpenMemHandle(&q0, nb.q[0], cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess)
            return fail(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        ctx->peer_opened.push_back(q0);
        if ((e = cudaIpcOpenMemHandle(&q1, nb.q[1], cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess)
            return fail(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        ctx->peer_opened.push_back(q1);
        if ((e = cudaIpcOpenMemHandle(&fl, nb.flags, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess)
            return fail(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        ctx->peer_opened.push_back(fl);
        pt.q[0][i] = static_cast<double*>(q0);
        pt.q[1][i] = static_cast<double*>(q1);
        pt.flag[i] = static_cast<unsigned long long*>(fl) + (mine >= 0 ? mine : 0);
        for (int k = 0; k < h.send_count; k++) {
            dst[(size_t)h.send_off + k] = nb.nb_recv_first[mine] + k;  // ghosts and send lists share the order (ascending global id)
            slot[(size_t)h.send_off + k] = (uint8_t)i;
        }
    }
    int r;
    if ((r = upload(ctx, &ctx->push_dst, dst))) return r;
    if ((r = upload(ctx, &ctx->push_slot, slot))) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->peer = pt;
    ctx->peer_ok = true;
    // graphs captured before the switch hold the NCCL exchange
    for (auto& ge : ctx->step_graph) if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
    return MSTGPU_OK;
}

// back to the NCCL exchange (a rank that could not connect makes every rank call this: the choice is collective)
int mstgpu_peer_disable(mstgpu_ctx* ctx) {
    if (!ctx) return MSTGPU_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream2));
    ctx->peer_ok = false;
    for (auto& ge : ctx->step_graph) if (ge) { cudaGraphExecDestroy(ge); ge = nullptr; }
    return MSTGPU_OK;
}

int mstgpu_mesh_adjacency(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int32_t* rowptr, int32_t* col, int64_t col_cap) {
    if (!mesh || !cfg || !rowptr) { set_error(nullptr, "null argument"); return MSTGPU_ERR_ARG; }
    Plan p;
    { std::string verr = validate_mesh(*mesh); if (!verr.empty()) { set_error(nullptr, verr); return MSTGPU_ERR_ARG; } }
    p.frame = choose_curve_frame(*mesh, *cfg);
    std::string perr = build_plan(*mesh, *cfg, p);
    if (!perr.empty()) { set_error(nullptr, perr); return MSTGPU_ERR_ARG; }
    const int nc = p.nc;
    rowptr[0] = 0;
    for (int c = 0; c < nc; c++) {
        int cnt = 1;
        for (int j = 0; j < p.nslot; j++) {
            const int v = p.cf[(size_t)j * nc + c];
            if (v >= 0 && ((v & 1) ? p.fc0[v >> 1] : p.fc1[v >> 1]) >= 0) cnt++;
        }
        const int64_t next = (int64_t)rowptr[c] + cnt;
        if (next > 0x7fffffffLL) { set_error(nullptr, "adjacency does not fit 32-bit offsets"); return MSTGPU_ERR_ARG; }
        rowptr[c + 1] = (int32_t)next;
    }
    if (!col) return MSTGPU_OK;  // sizing call
    if (col_cap < rowptr[nc]) { set_error(nullptr, "col buffer too small"); return MSTGPU_ERR_ARG; }
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nc; c++) {
        int32_t* out = col + rowptr[c];
        int n = 0;
        out[n++] = c;
        for (int j = 0; j < p.nslot; j++) {
            const int v = p.cf[(size_t)j * nc + c];
            if (v < 0) continue;
            const int nb = (v & 1) ? p.fc0[v >> 1] : p.fc1[v >> 1];
            if (nb >= 0) out[n++] = nb;
        }
        std::sort(out, out + n);
    }
    return MSTGPU_OK;
}

int mstgpu_plan_permutation(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int32_t* cell_new2old,
                            int32_t* face_new2old) {
    if (!mesh || !cfg) { set_error(nullptr, "null argument"); return MSTGPU_ERR_ARG; }
    Plan p;
    { std::string verr = validate_mesh(*mesh); if (!verr.empty()) { set_error(nullptr, verr); return MSTGPU_ERR_ARG; } }
    p.frame = choose_curve_frame(*mesh, *cfg);
    std::string perr = build_plan(*mesh, *cfg, p);
    if (!perr.empty()) { set_error(nullptr, perr); return MSTGPU_ERR_ARG; }
    if (cell_new2old) std::memcpy(cell_new2old, p.cell_new2old.data(), sizeof(int32_t) * p.nc);
    if (face_new2old) std::memcpy(face_new2old, p.face_new2old.data(), sizeof(int32_t) * p.nf);
    return MSTGPU_OK;
}

}  // extern "C"

#include "output.cuh"
