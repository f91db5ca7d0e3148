// GpuRhoSolver.h -- host-side mirror of the reference's density-based solver
// class, backed by the C ABI of include/mstgpu.h.
//
// The reference selects its solver by macro in Time::goNextTimeStep
// (R/time/Time.cpp:57-61) and drives it through a seven-method duck type
// (R/rhoSolver/RhoSolver.h:17-24; PSolver has the same shape):
//
//     RhoSolver(MshBlock*, fstream*, AllData*);  void setDT(NUM);  void solve();
//     VCTDIMU* getOldValue();  VCTDIMU* getNewValue();  VCTDIMU* getOldNTimeValue();
//     void updateNewToOld();
//
// GpuRhoSolverT has exactly these methods with the same meaning, so the only
// change in the reference is the type name at Time.cpp:58.
// (Identifiers avoid the reference's macro names -- DIM, DIMU, NUM, CV ... -- on purpose.)
// It is a template over the mesh / data classes only so that this header
// compiles without the reference tree; inside the reference it is used as
//     typedef GpuRhoSolverT<MshBlock, AllData, Face, Cell, VCTDIMU, DIM> GpuRhoSolver;
//
// Ownership mirrors the reference: Work owns MshBlock and
// AllData, the solver object is rebuilt on the stack EVERY step.  The heavy
// state (device tables, uploaded once) therefore lives in a context keyed by
// the (mesh, data) pair that outlives the per-step solver objects.
//
// Host <-> device traffic is lazy: solve() leaves the state on the device;
// getNewValue()/getOldValue() download into AllData's own arrays only when the
// host actually reads them (the reference's residual loop does, every step;
// a host that calls residual() instead avoids both copies).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mstgpu.h"

namespace mstgpu_host {

// The reference picks its scheme with macros at compile time (CONST.h:6 ACCURACY, :10 RHOSOLVER, :14 FLAGVISCID;
// shipped: first order, SolverAusm, inviscid).  A host that is compiled with the reference's CONST.h -- the drop-in
// case -- inherits exactly those; without the macros the defaults are the library's (second-order Roe, the
// configuration BASELINE.json quotes its metric on).
#define MSTGPU_HOST_STR_(x) #x
#define MSTGPU_HOST_STR(x) MSTGPU_HOST_STR_(x)
inline int flux_of_rhosolver_macro(const char* name) {
    for (const char* p = name; *p; p++)
        if ((p[0] == 'A' || p[0] == 'a') && (p[1] == 'u' || p[1] == 'U') && (p[2] == 's' || p[2] == 'S') && (p[3] == 'm' || p[3] == 'M'))
            return MSTGPU_FLUX_AUSM;
    return MSTGPU_FLUX_ROE;
}

struct Options {
#ifdef ACCURACY
    int order = (ACCURACY);        // ACCURACY      (R/include/CONST.h:6)
#else
    int order = 2;
#endif
#ifdef RHOSOLVER
    int flux = flux_of_rhosolver_macro(MSTGPU_HOST_STR(RHOSOLVER));  // RHOSOLVER (CONST.h:10)
#else
    int flux = MSTGPU_FLUX_ROE;
#endif
#ifdef FLAGVISCID
    int viscous = (FLAGVISCID);    // FLAGVISCID    (CONST.h:14)
#else
    int viscous = 0;
#endif
    int device = -1;
    double inletQ[5] = {0, 0, 0, 0, 0};
    bool have_inlet = false;
    // options the reference does not have; the defaults are the reference's scheme
    int gradient = MSTGPU_GRAD_GREEN_GAUSS;  // or MSTGPU_GRAD_LSQ
    int limiter = MSTGPU_LIMITER_NONE;       // or _BARTH_JESPERSEN / _VENKATAKRISHNAN
    double limiter_k = 5.0;
    // host_mirror: the fields live in AllData like in the reference -- solve() reads AllData's old array and writes
    // its new array (one mstgpu_step_host call: H2D, step and D2H pipelined over chunks of rows), the getters
    // return AllData's arrays without a copy.  For hosts that touch the cell state between steps; the default
    // keeps the state on the device and downloads lazily.
    bool host_mirror = false;
    bool implicit = false;                   // solve() = residual + block assembly + LU-SGS sweeps of R/lusolver
    int lusgs_iterations = 5;                // LU_INTERVAL (CONST.h:58)
};

// Flatten the reference's pointer-graph mesh into the tables of mstgpu_mesh,
// through the public getters only (R/mesh/{Face,Cell,FacesInf,MshBlock}.h).
template <class Mesh>
struct FlatMesh {
    int dim = 0;
    std::vector<int32_t> c0, c1, ftype, cf_ptr, cf_idx;
    std::vector<double> S, fc, eta, cc, vol;
    std::vector<int8_t> dac;
    std::vector<uint8_t> flag;
    mstgpu_mesh m{};

    void build(Mesh* mesh, int ND) {
        dim = ND;
        const int nc = mesh->getNumOfCells(), nf = mesh->getNumOfFaces();
        auto* faces = mesh->getBeginItFacesList();
        auto* cells = mesh->getBeginItCellsList();
        c0.resize(nf); c1.resize(nf); ftype.assign(nf, 0); eta.resize(nf); dac.resize(nf);
        S.resize((size_t)nf * ND); fc.resize((size_t)nf * ND); flag.resize((size_t)nf * ND);
        for (int f = 0; f < nf; f++) {
            for (int d = 0; d < ND; d++) {
                S[(size_t)f * ND + d] = faces[f].getDirect()[d];
                fc[(size_t)f * ND + d] = faces[f].getCenter()[d];
                flag[(size_t)f * ND + d] = faces[f].getFlagLeftRight()[d] ? 1 : 0;
            }
            eta[f] = faces[f].getEta0();
            dac[f] = (int8_t)faces[f].getDirectAndCells();
            c0[f] = faces[f].getBeginItPNbCells()[0]->getId();
            c1[f] = faces[f].getNumOfpNbCells() == 2 ? faces[f].getBeginItPNbCells()[1]->getId() : -1;
        }
        auto zones = mesh->getBeginItFacesInfList();
        for (int z = 0; z < mesh->getNumOfFacesInfs(); z++)
            for (int f = zones[z].getStart() - 1; f < zones[z].getEnd(); f++) ftype[f] = zones[z].getType();  // start is 1-based
        cc.resize((size_t)nc * ND); vol.resize(nc); cf_ptr.assign(nc + 1, 0);
        for (int c = 0; c < nc; c++) {
            vol[c] = cells[c].getVolume();
            for (int d = 0; d < ND; d++) cc[(size_t)c * ND + d] = cells[c].getCenter()[d];
            for (int j = 0; j < cells[c].getNumOfNbFaces(); j++) cf_idx.push_back(cells[c].getBeginItPNbFaces()[j]->getId());
            cf_ptr[c + 1] = (int32_t)cf_idx.size();
        }
        m.dim = ND; m.ncells = nc; m.nfaces = nf; m.nint = mesh->getNumOfIntFaces();
        m.c0 = c0.data(); m.c1 = c1.data(); m.S = S.data(); m.dac = dac.data(); m.fc = fc.data(); m.eta = eta.data();
        m.flag = flag.data(); m.ftype = ftype.data(); m.cc = cc.data(); m.vol = vol.data();
        m.cf_ptr = cf_ptr.data(); m.cf_idx = cf_idx.data();
    }
};

// Device context shared by the per-step solver objects of one (mesh, data) pair.
struct SharedContext {
    mstgpu_ctx* ctx = nullptr;
    int ncells = 0, U = 0;
    bool state_on_device = false;  // device holds the current state
    bool new_on_host = false, old_on_host = false;
    bool output_ready = false;     // mstgpu_output_setup done (nodeFields)
    std::vector<void*> pinned;     // AllData arrays page-locked for the streamed step (host_mirror)
    ~SharedContext() {
        for (void* q : pinned) mstgpu_host_unregister(q);
        if (ctx) mstgpu_destroy(ctx);
    }
};

inline std::map<std::pair<const void*, const void*>, SharedContext>& registry() {
    static std::map<std::pair<const void*, const void*>, SharedContext> r;
    return r;
}

inline void check(int rc, mstgpu_ctx* ctx, const char* what) {
    if (rc != MSTGPU_OK) throw std::runtime_error(std::string(what) + ": " + mstgpu_last_error(ctx));
}

template <class Mesh, class Data, class Face, class Cell, class Vec, int ND = 2>
class GpuRhoSolverT {
public:
    static Options& options() { static Options o; return o; }

    // RhoSolver::RhoSolver (R/rhoSolver/RhoSolver.cpp:3-26)
    GpuRhoSolverT(Mesh* mesh, std::fstream* flog, Data* data) : pMesh(mesh), pFlog(flog), pAllData(data) {
        sc = &registry()[{(const void*)mesh, (const void*)data}];
        if (!sc->ctx) {
            FlatMesh<Mesh> fm;
            fm.build(mesh, ND);
            mstgpu_config cfg;
            mstgpu_default_config(&cfg, ND);
            const Options& o = options();
            cfg.order = o.order; cfg.flux = o.flux; cfg.viscous = o.viscous; cfg.device = o.device;
            cfg.gradient = o.gradient; cfg.limiter = o.limiter; cfg.limiter_k = o.limiter_k;
            if (o.have_inlet) for (int k = 0; k < 5; k++) cfg.inletQ[k] = o.inletQ[k];
            check(mstgpu_create(&sc->ctx, &fm.m, &cfg), nullptr, "mstgpu_create");
            sc->ncells = fm.m.ncells; sc->U = ND + 2;
        }
    }
    void setDT(double dt) { DT = dt; }  // RhoSolver.cpp:33-35

    // RhoSolver::solve (RhoSolver.cpp:37-89).  The first call (or any call after the
    // host wrote AllData's old array, e.g. a restart) uploads the state.
    void solve() {
        if (options().host_mirror && !options().implicit) {
            // the reference's data flow: old array in, new array out, both on the host (RhoSolver.cpp:37-68)
            double* qo = reinterpret_cast<double*>(pAllData->getP1OldCellQs());
            double* qn = reinterpret_cast<double*>(pAllData->getP1NewCellQs());
            if (sc->pinned.empty()) {
                // page-locked buffers let the copies run beside the step; a failure only costs the overlap
                const size_t bytes = sizeof(double) * (size_t)sc->ncells * sc->U;
                for (double* q : {qo, qn})
                    if (mstgpu_host_register(q, bytes) == MSTGPU_OK) sc->pinned.push_back(q);
                if (sc->pinned.empty()) sc->pinned.push_back(nullptr);  // do not try again
            }
            check(mstgpu_step_host(sc->ctx, qo, qn, DT, 0), sc->ctx, "mstgpu_step_host");
            sc->state_on_device = sc->new_on_host = sc->old_on_host = true;
            return;
        }
        if (!sc->state_on_device) upload_old();
        if (options().implicit)
            check(mstgpu_step_implicit(sc->ctx, DT, 1, options().lusgs_iterations, nullptr), sc->ctx, "mstgpu_step_implicit");
        else
            check(mstgpu_step(sc->ctx, DT, 1), sc->ctx, "mstgpu_step");
        sc->new_on_host = sc->old_on_host = false;
    }
    Vec* getNewValue() {  // RhoSolver.cpp:507-509
        if (sc->state_on_device && !sc->new_on_host) {
            check(mstgpu_get_state(sc->ctx, reinterpret_cast<double*>(pAllData->getP1NewCellQs())), sc->ctx, "mstgpu_get_state");
            sc->new_on_host = true;
        }
        return pAllData->getP1NewCellQs();
    }
    Vec* getOldValue() {  // RhoSolver.cpp:504-506
        if (sc->state_on_device && !sc->old_on_host) {
            check(mstgpu_get_prev_state(sc->ctx, reinterpret_cast<double*>(pAllData->getP1OldCellQs())), sc->ctx, "mstgpu_get_prev_state");
            sc->old_on_host = true;
        }
        return pAllData->getP1OldCellQs();
    }
    Vec* getOldNTimeValue() { return pAllData->getPtP1OldNTimeCellQs(); }  // RhoSolver.cpp:510-512 (pseudo time is off)

    // RhoSolver::updateNewToOld (RhoSolver.cpp:513-517): a pointer swap on the device
    // (already done by mstgpu_step).  The host copy of "old" is refreshed because the
    // reference's output code reads getP1OldCellQs() (R/work/Work.cpp:42,67).
    void updateNewToOld() {
        if (sc->new_on_host)
            std::memcpy(pAllData->getP1OldCellQs(), pAllData->getP1NewCellQs(), sizeof(double) * (size_t)sc->ncells * sc->U);
        else
            check(mstgpu_get_state(sc->ctx, reinterpret_cast<double*>(pAllData->getP1OldCellQs())), sc->ctx, "mstgpu_get_state");
        sc->old_on_host = true;
    }

    // extras a GPU-aware host can use instead of reading both arrays every step
    void residual(double* out_dimu) { check(mstgpu_residual_linf(sc->ctx, out_dimu), sc->ctx, "mstgpu_residual_linf"); }
    void invalidateDeviceState() { sc->state_on_device = false; }  // call after writing AllData by hand

    // Output path: rho, u_i, T, p, Ma of every node for the current state, [nnodes][ND + 4], computed on the
    // device -- the numbers Work::writedataRhoBasedMshNodePlt prints after the coordinates
    // (R/work/Work.cpp:243-304), bit-identical, without downloading the cell state.  The node -> faces lists
    // and the per-node weight 1 / area(face[node id]) (Work.cpp:292) are taken from the host's own mesh once.
    void nodeFields(std::vector<double>& out) {
        if (!sc->state_on_device) upload_old();
        const int nn = pMesh->getNumOfNodes();
        if (!sc->output_ready) {
            FlatMesh<Mesh> fm;
            fm.build(pMesh, ND);
            auto* nodes = pMesh->getBeginItNodesList();
            auto* faces = pMesh->getBeginItFacesList();
            std::vector<int32_t> ptr((size_t)nn + 1, 0), idx;
            std::vector<double> w((size_t)nn);
            for (int i = 0; i < nn; i++) {
                for (auto it = nodes[i].getBeginItPNbFaces(); it < nodes[i].getEndItPNbFaces(); it++) idx.push_back((*it)->getId());
                ptr[(size_t)i + 1] = (int32_t)idx.size();
                w[(size_t)i] = 1 / faces[i].getArea();
            }
            check(mstgpu_output_setup(sc->ctx, &fm.m, nn, ptr.data(), idx.data(), w.data()), sc->ctx, "mstgpu_output_setup");
            sc->output_ready = true;
        }
        out.resize((size_t)nn * (ND + 4));
        check(mstgpu_node_fields(sc->ctx, out.data()), sc->ctx, "mstgpu_node_fields");
    }
    static void release(Mesh* mesh, Data* data) { registry().erase({(const void*)mesh, (const void*)data}); }

private:
    void upload_old() {
        static_assert(sizeof(Vec) == sizeof(double) * (ND + 2), "VCTDIMU must be DIMU plain doubles");
        check(mstgpu_set_state(sc->ctx, reinterpret_cast<const double*>(pAllData->getP1OldCellQs()), sc->ncells), sc->ctx, "mstgpu_set_state");
        sc->state_on_device = true;
    }
    Mesh* pMesh;
    std::fstream* pFlog;
    Data* pAllData;
    SharedContext* sc;
    double DT = 0.0;
};

}  // namespace mstgpu_host
