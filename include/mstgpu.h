/* mstgpu.h -- C ABI of the B200-native rhoSolver hot path.
 *
 * This is the drop-in boundary for MST-CFD's density-based solver
 * (R = /root/reference/MST-CFD).  Host code -- the reference's own .msh
 * reader, mesh classes and Time/Work driver -- stays as it is and reaches the
 * GPU through these entry points only: plain pointers and sizes, no C++ or
 * torch types.  What each entry point replaces:
 *
 *   mstgpu_create / mstgpu_destroy   RhoSolver::RhoSolver / ~RhoSolver
 *                                    (R/rhoSolver/RhoSolver.cpp:3-31); the
 *                                    reference rebuilds the solver every step
 *                                    (R/time/Time.cpp:58), the context lives
 *                                    outside it and is built once.
 *   mstgpu_mesh                      what the kernels read through
 *                                    Face::{getDirect,getCenter,getEta0,
 *                                    getFlagLeftRight,getBeginItPNbCells}
 *                                    (R/mesh/Face.cpp:94-133),
 *                                    Cell::{getVolume,getCenter,
 *                                    getBeginItDirectOfNbFaces,
 *                                    getBeginItPNbFaces} (R/mesh/Cell.cpp:69-137)
 *                                    and FacesInf::{getType,getStart,getEnd}
 *                                    (R/mesh/FacesInf.cpp:24-32), flattened
 *                                    once, in REFERENCE ORDER.
 *   mstgpu_config                    the compile-time macros of
 *                                    R/include/CONST.h as a runtime struct.
 *   mstgpu_set_state                 AllData::getP1OldCellQs() contents
 *                                    (R/data/AllData.cpp:3-27), host -> device.
 *   mstgpu_step                      RhoSolver::setDT + solve + updateNewToOld
 *                                    (RhoSolver.cpp:33-89, 513-517) repeated
 *                                    nsteps times = Time::goNextTimeStep
 *                                    (R/time/Time.cpp:54-81) without the host
 *                                    round trip.
 *   mstgpu_step_host                 one RhoSolver::solve with the fields where the
 *                                    reference keeps them: AllData's old array in,
 *                                    its new array out (Time.cpp:63-67), the copies
 *                                    overlapped with the step.
 *   mstgpu_residual_linf             the residual loop of Time.cpp:69-76.
 *   mstgpu_get_state                 RhoSolver::getNewValue (== old after
 *                                    updateNewToOld), device -> host.
 *   mstgpu_get_prev_state            RhoSolver::getOldValue as seen between
 *                                    solve() and updateNewToOld().
 *   mstgpu_debug_gradient            p1NewCellGradFlux   (RhoSolver.cpp:442-452)
 *   mstgpu_debug_face_flux           sum_d S[d] * p1OldFaceConvectFlux.col(d)
 *                                    per face (RhoSolver.cpp:53, 90-369)
 *   mstgpu_output_setup[_partitioned],
 *   mstgpu_node_fields               the arithmetic of
 *                                    Work::writedataRhoBasedMshNodePlt
 *                                    (R/work/Work.cpp:243-304) on the device.
 *   mstgpu_lusgs_*                   SparseSolverNUM::solveILUSGS
 *                                    (R/lusolver/SparseSolverNUM.cpp:144-212)
 *                                    and SparseSolver<MT,VCT>::solveILU
 *                                    (R/lusolver/SparseSolver.cpp:54-104).
 *
 * All state arrays are AoS [cell][DIMU] doubles in the reference's cell
 * numbering, exactly the memory image of VCTDIMU[] (Eigen fixed vectors are
 * plain doubles).  Return value: 0 = ok, negative = error (mstgpu_last_error
 * gives the text).  There is no CPU fallback: every entry point fails when no
 * CUDA device is usable.
 */
#ifndef MSTGPU_H
#define MSTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSTGPU_OK 0
#define MSTGPU_ERR_ARG -1
#define MSTGPU_ERR_CUDA -2
#define MSTGPU_ERR_STATE -3
#define MSTGPU_ERR_NCCL -4
#define MSTGPU_ERR_NAN -5

#define MSTGPU_FLUX_ROE 0  /* RHOSOLVER solverRoe  (R/rhoSolver/SolverRoe.cpp)  */
#define MSTGPU_FLUX_AUSM 1 /* RHOSOLVER SolverAusm (R/rhoSolver/SolverAusm.cpp) */

#define MSTGPU_GRAD_GREEN_GAUSS 0
#define MSTGPU_GRAD_LSQ 1
#define MSTGPU_LIMITER_NONE 0
#define MSTGPU_LIMITER_BARTH_JESPERSEN 1
#define MSTGPU_LIMITER_VENKATAKRISHNAN 2

/* zone types handled by RhoSolver::updateFaceFlux (RhoSolver.cpp:98-229) */
#define MSTGPU_BC_INTERIOR 2
#define MSTGPU_BC_WALL 3
#define MSTGPU_BC_OUTLET 5
#define MSTGPU_BC_SYMMETRY 7
#define MSTGPU_BC_INLET 10

typedef struct mstgpu_ctx mstgpu_ctx;
typedef struct mstgpu_part mstgpu_part;  /* one rank's piece of a mesh (mstgpu_partition_create, below) */

/* Flattened mesh, reference order.  D = dim, all arrays host memory. */
typedef struct mstgpu_mesh {
    int32_t dim;           /* DIM (2 or 3)                                          */
    int32_t ncells;        /* MshBlock::getNumOfCells                                */
    int32_t nfaces;        /* MshBlock::getNumOfFaces                                */
    int32_t nint;          /* MshBlock::getNumOfIntFaces                             */
    const int32_t* c0;     /* [nfaces] Face::getBeginItPNbCells()[0]->getId()        */
    const int32_t* c1;     /* [nfaces] ...[1]->getId(), -1 on boundary faces         */
    const double* S;       /* [nfaces*D] Face::getDirect() (area vector, as stored)  */
    const int8_t* dac;     /* [nfaces] Face::getDirectAndCells() (+1 / -1)           */
    const double* fc;      /* [nfaces*D] Face::getCenter()                           */
    const double* eta;     /* [nfaces] Face::getEta0()                               */
    const uint8_t* flag;   /* [nfaces*D] Face::getFlagLeftRight()                    */
    const int32_t* ftype;  /* [nfaces] zone type of the face (FacesInf::getType)     */
    const double* cc;      /* [ncells*D] Cell::getCenter()                           */
    const double* vol;     /* [ncells] Cell::getVolume()                             */
    const int32_t* cf_ptr; /* [ncells+1] CSR offsets of Cell::getBeginItPNbFaces()   */
    const int32_t* cf_idx; /* face ids per cell, in the cell's own (file) order      */
} mstgpu_mesh;

#define MSTGPU_TILE_DIRECT 1 /* packet words loaded straight into registers */
#define MSTGPU_TILE_STAGED 2 /* next face's weights / ids staged in shared memory by cp.async */

typedef struct mstgpu_config {
    int32_t order;        /* ACCURACY 1|2 (CONST.h:6)                                */
    int32_t flux;         /* MSTGPU_FLUX_* (CONST.h:10)                              */
    int32_t viscous;      /* FLAGVISCID (CONST.h:14)                                 */
    int32_t qf_copy_from; /* first face with Qf = Q[c0]; reference: nint-1
                             (the off-by-one of RhoSolver.cpp:438); <0 = nint-1     */
    int32_t renumber;     /* device cell order: 0 = reference order, 1 = Morton,
                             2 = Hilbert curve (default)                             */
    int32_t device;       /* CUDA ordinal, <0 = current device                       */
    double gamma;         /* GAMMA (CONST.h:41)                                      */
    double delta;         /* entropyError 0.125 (SolverRoe.cpp:115)                  */
    double eor;           /* EOR 1e-10 (CONST.h:42)                                  */
    double mu;            /* VISCIDMU (CONST.h:46)                                   */
    double kappa;         /* TEMPK (CONST.h:48)                                      */
    double cv;            /* CV (CONST.h:39)                                         */
    double inletQ[5];     /* inlet state (RhoSolver.cpp:123,266)                     */
    int32_t kernel;       /* 1 = fused tile kernel (default), 0 = three-kernel path;
                             viscous = 1 at FIRST order always runs the three-kernel path
                             (the fused kernel carries the viscous term at second order) */
    int32_t tile_cells;   /* cells per tile of the fused kernel, 0 = default         */
    int32_t block_threads;/* CTA size of the fused kernel (128|256), 0 = default     */
    int32_t tile_flags;   /* fused kernel, tets at second order: MSTGPU_TILE_DIRECT or
                             MSTGPU_TILE_STAGED; 0 = the library's default               */
    /* ---- build-defined extension: named by the project's north star, ABSENT from the
     * reference (no limiter, Green-Gauss only, fixed DT: SURVEY.md fact 2, 8f.4).  The
     * defaults (0, 0) are the reference's scheme.  Checked against the oracle's own
     * restatement of the same formulas ("parity unpinned": there is no reference code). */
    int32_t gradient;     /* MSTGPU_GRAD_*: 0 = Green-Gauss (RhoSolver.cpp:430-452),
                             1 = inverse-distance weighted least squares               */
    int32_t limiter;      /* MSTGPU_LIMITER_*: 0 = none, 1 = Barth-Jespersen,
                             2 = Venkatakrishnan (on the conserved variables)          */
    double limiter_k;     /* Venkatakrishnan K: eps^2 = (K h)^3, h = V^(1/D)            */
    int32_t tile_fit;     /* fused kernel: > 0 = tiles of VARIABLE size, each grown along the cell order until
                             its flux faces would exceed tile_fit (or its cells tile_cells): with tile_fit a
                             multiple of the CTA size every warp makes the same number of trips through the
                             flux phase.  0 = library default (512 = 4 x 128 threads for tets at second order,
                             768 = 3 x 256 for triangles at first order, when tile_cells and block_threads are 0
                             too; fixed tile_cells otherwise), < 0 = fixed tile_cells                        */
    int32_t reserved_;    /* keeps the struct a multiple of 8 bytes; set to 0            */
} mstgpu_config;

/* Fill `cfg` for `dim`: gas and flux constants are the reference's shipped ones (CONST.h:38-48,
 * SolverRoe.cpp:115).  The SCHEME is not: the reference chooses it with macros at compile time and
 * ships ACCURACY 1 / RHOSOLVER SolverAusm (CONST.h:6,10); this library defaults to order = 2,
 * flux = MSTGPU_FLUX_ROE (the configuration the project's metric is quoted on).  A host that replaces
 * the reference's solver sets cfg.order / cfg.flux / cfg.viscous from its own macros -- host/GpuRhoSolver.h
 * does exactly that when it is compiled with the reference's CONST.h. */
void mstgpu_default_config(mstgpu_config* cfg, int32_t dim);

/* Build a solver context: renumber (Morton), lay the tables out in HBM,
 * upload once.  The mesh arrays are not referenced after return. */
int mstgpu_create(mstgpu_ctx** out, const mstgpu_mesh* mesh, const mstgpu_config* cfg);
void mstgpu_destroy(mstgpu_ctx* ctx);

/* Host <-> device state in reference cell order, AoS [ncells][DIMU]. */
int mstgpu_set_state(mstgpu_ctx* ctx, const double* q_aos, int64_t ncells);
int mstgpu_get_state(mstgpu_ctx* ctx, double* q_aos);
int mstgpu_get_prev_state(mstgpu_ctx* ctx, double* q_aos);

/* Advance nsteps explicit steps of size dt (state stays on the device). */
int mstgpu_step(mstgpu_ctx* ctx, double dt, int32_t nsteps);
/* Same, bracketed by CUDA events on the solver's own stream; *ms = elapsed. */
int mstgpu_step_timed(mstgpu_ctx* ctx, double dt, int32_t nsteps, float* ms);
/* One step with the state on the HOST on both sides: what Time::goNextTimeStep does with a solver
 * whose fields live in AllData (R/time/Time.cpp:63-76: solve(), then getOldValue() / getNewValue() are
 * read on the host), i.e. mstgpu_set_state + mstgpu_step(dt, 1) + mstgpu_get_state in one call.
 * q_in / q_out: [ncells][DIMU] in reference order (q_in == q_out is allowed).  The three stages are
 * PIPELINED over `nchunks` chunks of host rows (<= 0: default 64, MSTGPU_HOST_CHUNKS overrides):
 * chunk j goes host -> device while the tiles whose input chunk j-1 completed run and the rows they
 * finished go device -> host, both PCIe directions busy at once.  The schedule (which tiles can run
 * after which chunk, which rows are final after which launch) is derived from the mesh numbering on
 * first use; it is valid for any numbering and hides the copies to the extent the host cell order
 * follows the mesh.  Results are bit-identical to the three separate calls.  Page-locked host buffers
 * (cudaHostRegister / cudaMallocHost) are needed for the overlap; pageable ones work, serialised.
 * Afterwards the context holds q_out as current and q_in as previous state, and
 * mstgpu_residual_linf returns the residual of this step.  On a partitioned context (COLLECTIVE) the
 * tiles next to ghost cells run after the halo exchange, which follows the last chunk.
 * Fused kernel only (kernel = 1). */
int mstgpu_step_host(mstgpu_ctx* ctx, const double* q_in, double* q_out, double dt, int32_t nchunks);
/* Page-lock / release a host array (cudaHostRegister / cudaHostUnregister) for hosts that are not built
 * against the CUDA headers: AllData's new[]-allocated field arrays (R/data/AllData.cpp:3-27) become
 * eligible for the overlapped copies of mstgpu_step_host.  unregister(NULL) is a no-op. */
int mstgpu_host_register(void* p, size_t bytes);
int mstgpu_host_unregister(void* p);
/* Extension (the reference's DT is the macro 1/STEP_TIME, R/time/Time.cpp:62):
 * global CFL time step  dt = cfl * min_c V_c / sum_{f in c} (|u_c.S_f| + a_c |S_f|)
 * of the current state.  With a communicator the minimum is taken over all ranks
 * (ncclAllReduce(min)): COLLECTIVE. */
int mstgpu_cfl_dt(mstgpu_ctx* ctx, double cfl, double* dt);
/* nsteps steps, each at the CFL step of its own start state; dt never visits the
 * host.  *time_advanced (optional) = sum of the steps taken.  Collective. */
int mstgpu_step_cfl(mstgpu_ctx* ctx, double cfl, int32_t nsteps, double* time_advanced);
/* Same, bracketed by CUDA events on the solver's own stream; *ms = elapsed (ms may be NULL). */
int mstgpu_step_cfl_timed(mstgpu_ctx* ctx, double cfl, int32_t nsteps, double* time_advanced, float* ms);
/* ---- implicit step: the LU-SGS sweeps of the reference's lusolver applied to rhoSolver -------------
 * The reference ships the solver (R/lusolver/SparseSolver.cpp:54-104) but never calls it from
 * rhoSolver, so the operator is build-defined: linearised backward Euler with
 * first-order flux Jacobians,
 *   [V_i/dt I + sum_f 1/2 (A(Q_i,S) + lam_f I)] dQ_i + sum_f 1/2 (A(Q_j,S) - lam_f I) dQ_j = -R_i(Q),
 * R = the explicit residual of mstgpu_step (same fluxes, same gather), solved by `lusgs_iters`
 * sweeps of the reference's block algorithm from dQ = 0; Q += dQ.  dt -> 0 recovers mstgpu_step.
 * setup builds the block pattern once (colour_sweeps != 0: colour-ordered sweeps, one level per
 * colour; 0: storage order, one level per wavefront); step_implicit calls it with colours if needed.
 * sweep_order: [ncells] reference cell ids in sweep order (for checking against the reference solver
 * on the permuted system).  *ms (optional) = CUDA-event time of the call. */
int mstgpu_implicit_setup(mstgpu_ctx* ctx, int32_t colour_sweeps);
int mstgpu_implicit_sweep_order(mstgpu_ctx* ctx, int32_t* order_ref_ids);
int mstgpu_step_implicit(mstgpu_ctx* ctx, double dt, int32_t nsteps, int32_t lusgs_iters, float* ms);
/* L-inf relative change of the LAST step, DIMU doubles (Time.cpp:69-76). */
int mstgpu_residual_linf(mstgpu_ctx* ctx, double* out_dimu);
int mstgpu_sync(mstgpu_ctx* ctx);

/* ---- output path: replaces the arithmetic of
 * Work::writedataRhoBasedMshNodePlt (R/work/Work.cpp:243-304), which the reference runs on the
 * host after reading the whole cell state back.
 *   mstgpu_output_setup  once: the node -> faces lists in the order Node::addNbFace builds them
 *                        (R/mesh/Node.cpp:13-15; nf_ptr [nnodes+1], nf_idx), the mesh the context was
 *                        created from (c0, c1, eta, ftype are read), and the per-node weight of
 *                        Work.cpp:292-293, 1 / Face::getArea() of face[node id] (NULL = 1).
 *   mstgpu_node_fields   out [nnodes][D+4] = rho, u_i, T, p, Ma of every node for the CURRENT
 *                        state, bit-identical to the reference's numbers (the columns it prints
 *                        after the coordinates, Work.cpp:299-303).  Zone types its switch does not
 *                        list (symmetry) read an uninitialised array there; Q[c0] here.
 * A partitioned context takes mstgpu_output_setup_partitioned instead: the GLOBAL mesh, node lists and weights
 * (what every rank has read anyway) and the partition the context was created from.  Every node is computed by
 * exactly one rank -- the owner of c0 of the first face in its list -- from fresh rows: the states of other ranks'
 * cells around its nodes (they can lie beyond the ghost layers of the step) are fetched inside mstgpu_node_fields,
 * which is then COLLECTIVE (one pack kernel + grouped ncclSend / ncclRecv; lists derived on every rank from the
 * global tables, no negotiation).  mstgpu_output_node_count / _ids give the nodes of this rank's rows (global ids,
 * ascending; all nodes in order on an unpartitioned context); out is [count][D+4].  Bit-identical to the
 * single-GPU numbers (same faces, same order, same arithmetic per node). */
int mstgpu_output_setup(mstgpu_ctx* ctx, const mstgpu_mesh* mesh, int32_t nnodes, const int32_t* nf_ptr,
                        const int32_t* nf_idx, const double* node_weight);
int mstgpu_output_setup_partitioned(mstgpu_ctx* ctx, const mstgpu_part* part, const mstgpu_mesh* global_mesh, int32_t nnodes,
                                    const int32_t* nf_ptr, const int32_t* nf_idx, const double* node_weight);
int32_t mstgpu_output_node_count(mstgpu_ctx* ctx);
int mstgpu_output_node_ids(mstgpu_ctx* ctx, int32_t* ids);
int mstgpu_node_fields(mstgpu_ctx* ctx, double* out);

/* Stage probes of the last step, reference order.
 * gradient: [ncells][DIMU][D]; face flux: [nfaces][DIMU] =
 * sum_d (dac*S)[d] * F[:,d], i.e. the flux through the face oriented out of c0. */
int mstgpu_debug_gradient(mstgpu_ctx* ctx, double* grad);
int mstgpu_debug_face_flux(mstgpu_ctx* ctx, double* phi);

/* Experimental: launch variant of the default fused instantiation (3-D, second order, 256 threads, no
 * extension): bit mask, 1 = L2 prefetch one tile ahead, 2 = persistent CTAs, 4 = cp.async ring rows;
 * 0 = the measured default.  2-D first order on triangles: 8 / 16 = registers sized for 4 / 3 resident
 * CTAs per SM (the default picks 4 for AUSM+, 3 for Roe), 32 = the plain 2-CTA allocation.  Same arithmetic
 * in the same order: results are bit-identical.  The environment variable MSTGPU_TILE_VAR sets it at
 * creation.  (csrc/step_tiles.cuh) */
int mstgpu_set_tile_variant(mstgpu_ctx* ctx, int32_t variant);

/* Introspection for the bench: kernels launched so far by this context and
 * per-kernel accumulated device time (ms) when timing is enabled. */
int64_t mstgpu_launch_count(mstgpu_ctx* ctx);
int mstgpu_enable_kernel_timing(mstgpu_ctx* ctx, int32_t on);
/* names: "gradient", "flux", "update"; returns accumulated ms and launches */
int mstgpu_kernel_time(mstgpu_ctx* ctx, const char* name, double* ms, int64_t* launches);
int64_t mstgpu_device_bytes(mstgpu_ctx* ctx);

/* ---- multi-GPU: one partition per GPU, ghost-cell halo exchange over NCCL --------
 * The reference has no distributed path; these entry points are
 * the build's own.  Every rank holds the global flat mesh (as the reference's
 * host code does after reading the .msh), asks for its partition, creates a
 * context on it and joins a communicator.  State moves in the partition's own
 * cell order: owned cells by ascending global id (mstgpu_partition_cell_ids).
 * Results are bit-identical to the single-GPU run for any partition count. */

/* cell_part: [ncells] partition of every global cell, or NULL for equal ranges of
 * the Hilbert curve.  Host only (no CUDA call). */
int mstgpu_partition_create(mstgpu_part** out, const mstgpu_mesh* global_mesh, const mstgpu_config* cfg,
                            int32_t nparts, int32_t rank, const int32_t* cell_part);
void mstgpu_partition_destroy(mstgpu_part* part);
/* local mesh: owned cells, then ghost cells grouped by owner rank */
const mstgpu_mesh* mstgpu_partition_mesh(const mstgpu_part* part);
int mstgpu_partition_sizes(const mstgpu_part* part, int32_t* n_owned, int32_t* n_local, int32_t* n_neighbors);
const int32_t* mstgpu_partition_cell_ids(const mstgpu_part* part); /* [n_local] local -> global */
int mstgpu_partition_neighbor(const mstgpu_part* part, int32_t i, int32_t* rank, int32_t* send_count,
                              const int32_t** send_local, int32_t* recv_first, int32_t* recv_count);

/* Context on a partition: set/get_state move the n_owned owned rows only. */
int mstgpu_create_partitioned(mstgpu_ctx** out, const mstgpu_part* part, const mstgpu_config* cfg);
/* NCCL bootstrap: rank 0 obtains an id, the host broadcasts the 128 bytes by its
 * own means (MPI, torch.distributed, a file), every rank calls comm_init. */
int mstgpu_comm_unique_id(char* out128);
int mstgpu_comm_init(mstgpu_ctx* ctx, int32_t nranks, int32_t rank, const char* id128);

/* Peer-memory halo (one process per GPU on one NVLink / NVSwitch node): instead of pack + ncclSend/ncclRecv,
 * each rank's own kernel STORES its boundary rows into the neighbours' ghost blocks through peer pointers and
 * publishes an epoch flag; receivers wait (bounded, ~10 s -> MSTGPU_ERR_NCCL) on their flags.  Set-up is an
 * exchange of fixed-size blobs (CUDA IPC handles + row offsets) that the host all-gathers by whatever means it
 * has: export on every rank, gather, connect on every rank (collective; after mstgpu_comm_init, which stays in
 * charge of the residual / time-step reductions).  If connect fails (no peer access between two devices) the
 * context keeps exchanging through NCCL.  Results are bit-identical either way. */
int64_t mstgpu_peer_blob_bytes(void);
int mstgpu_peer_export(mstgpu_ctx* ctx, int32_t rank, void* blob);
int mstgpu_peer_connect(mstgpu_ctx* ctx, int32_t nranks, int32_t rank, const void* blobs /* nranks blobs, by rank */);
int mstgpu_peer_disable(mstgpu_ctx* ctx); /* back to the NCCL exchange (collective, like connect) */
/* With a communicator, mstgpu_step exchanges ghost states before every step and
 * mstgpu_residual_linf is COLLECTIVE (max over ranks, ncclAllReduce). */

/* ---- LU-SGS sweeps of the reference's lusolver ------------------------------------
 * Replaces SparseSolverNUM::{setELE,addELE,setD,setRHSb,solveILUSGS,getPNewX}
 * (R/lusolver/SparseSolverNUM.h:13-35) for block = 1 and
 * SparseSolver<MT,VCT>::{setD,setL,setU,setRHSb,solveILU,getItBeginX}
 * (R/lusolver/SparseSolver.h:13-22) for block = DIMU (4 or 5).  The matrix is
 * handed over as CSR by row (columns ascending; blocks row-major, block x
 * block doubles per entry): the entries a host would pass one by one to
 * setELE / setD / setL / setU.  The sweeps run level by level and reproduce the
 * reference's sequential result on the matrix ordering given;
 * mstgpu_lusgs_color_order returns a colour ordering (few levels) for hosts that
 * want to permute their system first. */
typedef struct mstgpu_lusgs mstgpu_lusgs;
int mstgpu_lusgs_create(mstgpu_lusgs** out, int32_t n, int32_t block, const int32_t* rowptr, const int32_t* col,
                        int32_t device);
/* Same with a sweep order: sweep_new2old[i] = row visited i-th (a permutation, e.g. from
 * mstgpu_lusgs_color_order; NULL = storage order).  The result is that of the reference's solver on
 * the permuted system P A P^T, P b -- but the matrix, b and x stay in STORAGE order (locality of
 * the mesh numbering is kept; only the dependency levels and the order of a row's terms change). */
int mstgpu_lusgs_create_ordered(mstgpu_lusgs** out, int32_t n, int32_t block, const int32_t* rowptr,
                                const int32_t* col, const int32_t* sweep_new2old, int32_t device);
/* One partition of a distributed system: n rows this rank owns, columns in
 * [n, ncols) are rows other ranks own ("ghost columns").  Their couplings are LAGGED: every
 * iteration first forms b - sum_ghost A[r,c] x[c] with the x the caller holds for them (x has ncols
 * rows; the caller refreshes the ghost rows between iterations, e.g. one iteration per call) -- block
 * Jacobi across partitions, the reference's sweeps inside each. */
int mstgpu_lusgs_create_partitioned(mstgpu_lusgs** out, int32_t n, int32_t ncols, int32_t block, const int32_t* rowptr,
                                    const int32_t* col, const int32_t* sweep_new2old, int32_t device);
int mstgpu_lusgs_color_order_partitioned(int32_t n, int32_t ncols, const int32_t* rowptr, const int32_t* col,
                                         int32_t* perm_new2old, int32_t* ncolors);
void mstgpu_lusgs_destroy(mstgpu_lusgs* h);
/* x: in = start vector (the constructors' pOldX), out = solution.  max_iter = LU_INTERVAL
 * (CONST.h:58).  early_exit != 0 applies the scalar version's stop test
 * 1e-20 < res < 1e-7 (SparseSolverNUM.cpp:205).  res_hist: [max_iter] or NULL. */
int mstgpu_lusgs_solve(mstgpu_lusgs* h, const double* val, const double* b, double* x, int32_t max_iter,
                       int32_t early_exit, double* res_hist, int32_t* iters_done);
/* Device-resident solve: d_val / d_b / d_x are DEVICE pointers in the layout of mstgpu_lusgs_solve
 * (x in place); nothing crosses PCIe.  Runs max_iter iterations (the block version's fixed count).
 * *ms (optional) = CUDA-event time of the whole solve on the solver's stream. */
int mstgpu_lusgs_solve_device(mstgpu_lusgs* h, const double* d_val, const double* d_b, double* d_x, int32_t max_iter,
                              float* ms);
/* Iteration variants (csrc/lusgs.cu), all the mathematics of SparseSolver.cpp:54-104:
 *   0  the reference's passes one by one (U x, right-hand side, forward sweep, D D^-1, backward sweep, D^-1);
 *   1  fused: the backward sweep leaves U x of the next iteration as a by-product and the right-hand side update is
 *      folded into the forward sweep -- the unscaled off-diagonal blocks are read once per solve instead of twice
 *      per iteration (re-associated: 6e-16 against the reference-pinned oracle, tests/lusgs_fused_np.py);
 *   2  lean (default): mode 1 without the reference's w0 = D (D^-1 v) round trip (the identity up to cond(D) eps)
 *      and with x = D^-1 w formed inside the backward sweep (<= 1e-13 against the oracle on the test systems).
 * The scalar solver's residual history / early exit needs the last pass of modes 0 / 1 and gets it in mode 2 too.
 * MSTGPU_LUSGS_MODE sets the mode at creation. */
int mstgpu_lusgs_set_mode(mstgpu_lusgs* h, int32_t mode);
int64_t mstgpu_lusgs_launch_count(mstgpu_lusgs* h);
int64_t mstgpu_lusgs_device_bytes(mstgpu_lusgs* h);
int mstgpu_lusgs_levels(mstgpu_lusgs* h, int32_t* forward_levels, int32_t* backward_levels);
/* host only: greedy colouring of the symmetrised pattern; rows sorted by colour */
int mstgpu_lusgs_color_order(int32_t n, const int32_t* rowptr, const int32_t* col, int32_t* perm_new2old,
                             int32_t* ncolors);
const char* mstgpu_lusgs_last_error(void);

/* Host-only: the renumbering mstgpu_create would apply (no CUDA call), for
 * inspection and CPU tests.  cell_new2old [ncells], face_new2old [nfaces]. */
int mstgpu_plan_permutation(const mstgpu_mesh* mesh, const mstgpu_config* cfg,
                            int32_t* cell_new2old, int32_t* face_new2old);

/* Host-only: the pattern of an implicit operator on this mesh -- CSR adjacency through interior
 * faces plus the diagonal, columns ascending, cells in the DEVICE order mstgpu_create would use
 * (mstgpu_plan_permutation) -- ready for mstgpu_lusgs_create[_ordered].  rowptr [ncells+1];
 * col may be NULL for a sizing call (rowptr[ncells] entries are needed), col_cap = its capacity. */
int mstgpu_mesh_adjacency(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int32_t* rowptr, int32_t* col,
                          int64_t col_cap);

/* Host-only: statistics of the tiling the fused kernel would use.
 * out[0]=tiles out[1]=max smem bytes out[2]=mean smem bytes out[3]=sum ring1
 * out[4]=sum ring2 out[5]=sum flux faces out[6]=sum local faces out[7]=packet bytes
 * out[8..11]=smem histogram: tiles needing <=56K, <=75K, <=113K, more
 * out[12..14]=sum over tiles of the loop trips of a CTA: ceil(flux faces / NT), ceil(owned cells / NT),
 * ceil(ring cells / NT); out[15]=NT.  (16 entries.) */
int mstgpu_tile_stats(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int64_t* out16);
/* Same for a partition's local mesh (mstgpu_partition_mesh): cells [0, n_owned) are renumbered and tiled, the
 * ghost cells behind them only lend their state -- the tiling mstgpu_create_partitioned builds.  n_owned < 0 = all. */
int mstgpu_tile_stats_owned(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int32_t n_owned, int64_t* out);
/* Memory locality of the tiling (host only): out6 = tiles, ring rows, distinct 128-byte lines the ring rows of a tile
 * touch (summed over tiles), runs of consecutive ring ids, 128-byte lines requested by the ring gather (per warp load
 * instruction), flux faces.  Diagnostic for the cell order (profiles/r2_scaling.md). */
int mstgpu_tile_locality(const mstgpu_mesh* mesh, const mstgpu_config* cfg, int32_t n_owned, int64_t* out6);

const char* mstgpu_last_error(mstgpu_ctx* ctx); /* ctx may be NULL (create errors) */
const char* mstgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MSTGPU_H */
