// mstrun.cpp -- the reference's main program for the density-based solver, on the GPU, without the
// reference tree: what main.cpp + Work::work + Time do for RHO_P == 0 (R = /root/reference/MST-CFD),
// through the two C ABIs only (include/msthost.h, include/mstgpu.h).
//
//   R/main.cpp:4-11            mesh name                                   -> argv[1]
//   R/work/Work.cpp:23-25      log file result/<msh>_TIME4000_u0-log.lhblog -> <out>/<msh>_TIME4000_u0-log.lhblog
//   R/work/Work.cpp:26         MshBlock::readMsh                           -> msthost_msh_read + msthost_flatten
//   R/work/Work.cpp:31-38      Time::initialization (R/time/Time.cpp:13-52: base state of CONST.h:70-75,
//                              rho x 0.125 and E x 0.1 where the cell centre has x > 0.5)
//   R/work/Work.cpp:42,64-72   writedataRhoBasedMshNodePlt at t = 0 and every SAVE_TIME*STEP_TIME = 10 steps
//                                                                          -> mstgpu_node_fields + msthost_plt_write
//   R/time/Time.cpp:54-81      goNextTimeStep: setDT(1/STEP_TIME), solve, residual line in the log, new -> old
//                                                                          -> mstgpu_step + mstgpu_residual_linf
// Not reproduced: the stdin prompt every 3200 steps (Work.cpp:80-88; --steps bounds the run instead), the
// p-based restart dump (.cellPUVT, Work.cpp:73-76: it stores the pressure solver's array).
// The macros of R/include/CONST.h that select scheme and order are options here.
//
//   mstrun <mesh.msh> [--steps N=400] [--order 1|2] [--flux roe|ausm] [--dt 2.5e-4] [--save-every 10]
//          [--out DIR=result] [--flags consistent|as_shipped] [--init sod|uniform] [--binary] [--device K]
//
// No CPU fallback: without a usable CUDA device the program stops with the library's error text.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mstgpu.h"
#include "../../include/msthost.h"

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

[[noreturn]] void die(const std::string& what) {
    fprintf(stderr, "mstrun: %s\n", what.c_str());
    exit(1);
}

// `ostream << fixed << setprecision(15) << setw(15)` of Time.cpp:78 (setw binds to the first number only)
std::string fixed15(double v, bool setw15) {
    char t[400];
    const int n = snprintf(t, sizeof t, "%.15f", v);
    std::string s(t, (size_t)n);
    if (setw15 && n < 15) s.insert(0, (size_t)(15 - n), ' ');
    return s;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
        fprintf(stderr, "usage: mstrun mesh.msh [--steps N] [--order 1|2] [--flux roe|ausm] [--dt DT] [--save-every K]\n"
                        "              [--out DIR] [--flags consistent|as_shipped] [--init sod|uniform] [--binary] [--device K]\n");
        return argc < 2 ? 2 : 0;
    }
    std::string msh = argv[1], out = "result", flags = "consistent", init = "sod", flux = "roe";
    int steps = 400, order = 2, save_every = 10, device = -1;  // ACCURACY 2; SAVE_TIME * STEP_TIME = 10 (CONST.h:52-53)
    double dt = 1. / 4e+3;                                      // 1 / STEP_TIME (Time.cpp:62)
    bool binary = false;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) die("missing value after " + a); return argv[++i]; };
        if (a == "--steps") steps = atoi(val());
        else if (a == "--order") order = atoi(val());
        else if (a == "--flux") flux = val();
        else if (a == "--dt") dt = atof(val());
        else if (a == "--save-every") save_every = atoi(val());
        else if (a == "--out") out = val();
        else if (a == "--flags") flags = val();
        else if (a == "--init") init = val();
        else if (a == "--device") device = atoi(val());
        else if (a == "--binary") binary = true;
        else die("unknown option " + a);
    }
    if ((order != 1 && order != 2) || (flux != "roe" && flux != "ausm") || (flags != "consistent" && flags != "as_shipped") ||
        (init != "sod" && init != "uniform") || steps < 0 || save_every < 1)
        die("bad option value (see --help)");

    // ---- MshBlock::readMsh ------------------------------------------------------------------------
    double t0 = now_s();
    msthost_msh* h = nullptr;
    if (msthost_msh_read(msh.c_str(), &h) != 0) die(std::string("reading ") + msh + ": " + msthost_last_error());
    int64_t sz[8];
    msthost_msh_sizes(h, sz);
    const int D = (int)sz[0], U = D + 2, npf = (int)sz[6];
    const int64_t nn = sz[1], nc = sz[2], nf = sz[3], nint = sz[4];
    std::vector<double> nodes((size_t)(nn * D));
    std::vector<int32_t> face_nodes((size_t)(nf * npf)), c0((size_t)nf), c1((size_t)nf), ftype((size_t)nf);
    msthost_msh_tables(h, nodes.data(), face_nodes.data(), c0.data(), c1.data(), ftype.data(), nullptr);
    msthost_msh_free(h);
    std::vector<double> S((size_t)(nf * D)), fc((size_t)(nf * D)), eta((size_t)nf), cc((size_t)(nc * D)), vol((size_t)nc, 0.0);
    std::vector<int8_t> dac((size_t)nf);
    std::vector<uint8_t> flag((size_t)(nf * D));
    std::vector<int32_t> cf_ptr((size_t)nc + 1), cf_idx((size_t)(nf + nint));
    if (msthost_flatten(D, nn, nc, nf, npf, nodes.data(), face_nodes.data(), c0.data(), c1.data(), flags == "as_shipped" ? 1 : 0,
                        S.data(), fc.data(), dac.data(), eta.data(), flag.data(), cc.data(), vol.data(), cf_ptr.data(),
                        cf_idx.data()) != 0)
        die("msthost_flatten failed");
    std::vector<int32_t> nf_ptr((size_t)nn + 1), nf_idx;
    {
        int64_t pairs = 0;
        for (int32_t v : face_nodes) pairs += v >= 0;
        nf_idx.resize((size_t)pairs);
    }
    if (msthost_node_faces(nn, nf, npf, face_nodes.data(), nf_ptr.data(), nf_idx.data()) != 0) die("msthost_node_faces failed");
    std::vector<int32_t> cn_ptr((size_t)nc + 1), cn_idx;
    msthost_cell_nodes(D, nc, npf, face_nodes.data(), cf_ptr.data(), cf_idx.data(), nodes.data(), cn_ptr.data(), nullptr);
    cn_idx.resize((size_t)cn_ptr[(size_t)nc]);
    msthost_cell_nodes(D, nc, npf, face_nodes.data(), cf_ptr.data(), cf_idx.data(), nodes.data(), cn_ptr.data(), cn_idx.data());
    const double t_mesh = now_s() - t0;

    // ---- solver context (RhoSolver::RhoSolver, once) ------------------------------------------------
    t0 = now_s();
    mstgpu_mesh m{};
    m.dim = D; m.ncells = (int32_t)nc; m.nfaces = (int32_t)nf; m.nint = (int32_t)nint;
    m.c0 = c0.data(); m.c1 = c1.data(); m.S = S.data(); m.dac = dac.data(); m.fc = fc.data(); m.eta = eta.data();
    m.flag = flag.data(); m.ftype = ftype.data(); m.cc = cc.data(); m.vol = vol.data(); m.cf_ptr = cf_ptr.data(); m.cf_idx = cf_idx.data();
    mstgpu_config cfg;
    mstgpu_default_config(&cfg, D);
    cfg.order = order;
    cfg.flux = flux == "roe" ? MSTGPU_FLUX_ROE : MSTGPU_FLUX_AUSM;
    cfg.device = device;
    mstgpu_ctx* ctx = nullptr;
    if (mstgpu_create(&ctx, &m, &cfg) != MSTGPU_OK) die(std::string("mstgpu_create: ") + mstgpu_last_error(nullptr));
    auto check = [&](int rc, const char* what) { if (rc != MSTGPU_OK) die(std::string(what) + ": " + mstgpu_last_error(ctx)); };
    // node weights of Work.cpp:292: 1 / area(face[node id])
    if (nf < nn) die("fewer faces than nodes: the reference's writer indexes faces with node ids");
    std::vector<double> w((size_t)nn);
    for (int64_t i = 0; i < nn; i++) {
        double a2 = S[(size_t)(i * D)] * S[(size_t)(i * D)] + S[(size_t)(i * D + 1)] * S[(size_t)(i * D + 1)];
        if (D == 3) a2 = a2 + S[(size_t)(i * D + 2)] * S[(size_t)(i * D + 2)];
        w[(size_t)i] = 1 / std::sqrt(a2);
    }
    check(mstgpu_output_setup(ctx, &m, (int32_t)nn, nf_ptr.data(), nf_idx.data(), w.data()), "mstgpu_output_setup");

    // ---- Time::initialization (Time.cpp:13-38; CONST.h:70-75) ---------------------------------------
    std::vector<double> Q((size_t)(nc * U), 0.0);
    {
        const double iniT = 1 / 286.32, iniE = 1 * (iniT * 715.8 + 0.5 * (0 * 0 + 0 * 0));
        for (int64_t c = 0; c < nc; c++) {
            double rho = 1, E = iniE;
            if (init == "sod" && cc[(size_t)(c * D)] > 0.5) { rho *= 0.125; E *= 0.1; }
            Q[(size_t)(c * U)] = rho;
            Q[(size_t)(c * U + U - 1)] = E;
        }
    }
    check(mstgpu_set_state(ctx, Q.data(), nc), "mstgpu_set_state");
    const double t_setup = now_s() - t0;

    const std::string base = msh.substr(msh.find_last_of('/') + 1);
    const std::string stem = out + "/" + base + "_TIME4000_u0";  // Work.cpp:21: "_TIME" << STEP_TIME, "_u" << inletu
    FILE* flog = fopen((stem + "-log.lhblog").c_str(), "w");
    if (!flog) die("cannot create " + stem + "-log.lhblog (the directory must exist, like the reference's result/)");
    std::vector<double> fields((size_t)(nn * (D + 4)));
    double t_out = 0.0, t_step = 0.0;
    auto write_plt = [&](int t) {
        const double a = now_s();
        check(mstgpu_node_fields(ctx, fields.data()), "mstgpu_node_fields");
        const std::string path = stem + "_t" + std::to_string(t) + (binary ? ".pltbin" : ".plt");
        const int rc = binary ? msthost_plt_write_binary(path.c_str(), D, nn, nc, nodes.data(), fields.data(), cn_ptr.data(), cn_idx.data(), t)
                              : msthost_plt_write(path.c_str(), D, nn, nc, nodes.data(), fields.data(), cn_ptr.data(), cn_idx.data(), t, 4);
        if (rc != 0) die("cannot write " + path);
        t_out += now_s() - a;
    };
    write_plt(0);  // Work.cpp:42

    // ---- the step loop (Work.cpp:53-89, Time.cpp:54-81) ----------------------------------------------
    std::vector<double> r((size_t)U);
    for (int t = 0; t < steps; t++) {
        const double a = now_s();
        check(mstgpu_step(ctx, dt, 1), "mstgpu_step");
        const int rc = mstgpu_residual_linf(ctx, r.data());
        if (rc != MSTGPU_OK && rc != MSTGPU_ERR_NAN) check(rc, "mstgpu_residual_linf");
        t_step += now_s() - a;
        std::string line = fixed15(r[0], true);
        for (int k = 1; k < 4; k++) line += " " + fixed15(r[(size_t)k], false);  // the reference logs 4 components
        fprintf(flog, "%s \n", line.c_str());
        if ((t + 1) % save_every == 0) write_plt(t + 1);
    }
    fclose(flog);
    printf("mstrun: %lld cells, %lld faces, %lld nodes; mesh %.3f s, setup %.3f s, %d steps %.3f s (%.3f ms/step with the residual read back), "
           "%d output files %.3f s\n", (long long)nc, (long long)nf, (long long)nn, t_mesh, t_setup, steps, t_step,
           steps ? 1e3 * t_step / steps : 0.0, 1 + steps / save_every, t_out);
    mstgpu_destroy(ctx);
    return 0;
}
