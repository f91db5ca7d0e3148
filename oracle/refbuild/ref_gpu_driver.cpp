// INTEGRATION HARNESS (test infrastructure).  The reference's OWN host code --
// MshBlock .msh reader, AllData, Time::initialization -- driving the GPU solver
// through mst-cfd_b200/host/GpuRhoSolver.h, in the call sequence of
// Time::goNextTimeStep (R/time/Time.cpp:54-81).  Same command line and dump
// format as ref_driver.cpp, so tests compare "reference host + GPU solver" with
// "reference host + reference CPU solver" file against file.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "time/Time.h"
#include "GpuRhoSolver.h"

typedef mstgpu_host::GpuRhoSolverT<MshBlock, AllData, Face, Cell, VCTDIMU, DIM> GpuRhoSolver;

static FILE* g_out;
static void dump(const char* name, const char* dtype, const void* p, long long count, int width) {
    fprintf(g_out, "%s %s %lld\n", name, dtype, count);
    fwrite(p, width, (size_t)count, g_out);
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: ref_gpu_* mesh.msh out.bin flagmode retag init steps...\n"); return 2; }
    std::string msh = argv[1];
    g_out = fopen(argv[2], "wb");
    const int flagmode = atoi(argv[3]);
    std::string retag = argv[4], init = argv[5];
    std::vector<int> steps;
    for (int i = 6; i < argc; i++) steps.push_back(atoi(argv[i]));

    MshBlock mesh;
    mesh.readMsh(msh);
    const int nc = mesh.getNumOfCells(), nf = mesh.getNumOfFaces();
    Face* faces = mesh.getBeginItFacesList();
    if (retag != "-") {
        int a = atoi(retag.substr(0, retag.find(':')).c_str()), b = atoi(retag.substr(retag.find(':') + 1).c_str());
        auto it = mesh.getBeginItFacesInfList();
        for (int z = 0; z < mesh.getNumOfFacesInfs(); z++)
            if (it[z].getType() == a) it[z].setType(b);
    }
    if (flagmode == 1)
        for (int f = 0; f < nf; f++)
            for (int d = 0; d < DIM; d++)
                faces[f].setFlagLeftRight(d, faces[f].getDirectAndCells() * faces[f].getDirect()[d] >= 0);
    int hdr[4] = {DIM, nc, nf, mesh.getNumOfIntFaces()};
    dump("hdr", "i4", hdr, 4, 4);

    AllData allData;
    allData.createAllData(nc, nf);
    std::fstream fLog("/dev/null", std::ios::out);
    Time time1(&mesh, &fLog, &allData);
    VCTDIMU iniQ;
    iniQ << inirho, inirho * iniu, inirho * iniv, iniE;
    time1.initialization(iniQ);
    if (init != "-") {
        std::vector<double> q((size_t)nc * (DIMU));
        FILE* fi = fopen(init.c_str(), "rb");
        if (!fi || fread(q.data(), 8, q.size(), fi) != q.size()) { fprintf(stderr, "bad init file\n"); return 3; }
        fclose(fi);
        for (int c = 0; c < nc; c++)
            for (int k = 0; k < DIMU; k++) allData.getP1OldCellQs()[c][k] = allData.getP1NewCellQs()[c][k] = q[(size_t)c * (DIMU) + k];
    }
    GpuRhoSolver::options().order = ACCURACY;
    GpuRhoSolver::options().flux = MST_FLUX;
    // MST_HOST_MIRROR=1: the fields stay in AllData like in the reference, solve() = one streamed step
    // (mstgpu_step_host: old array in, new array out)
    if (const char* hm = getenv("MST_HOST_MIRROR")) GpuRhoSolver::options().host_mirror = atoi(hm) != 0;
    dump("Q0", "f8", allData.getP1OldCellQs(), (long long)nc * (DIMU), 8);
    std::vector<double> resid;
    int done = 0;
    for (int s : steps) {
        for (; done < s; done++) {
            // ---- Time::goNextTimeStep with the GPU solver (Time.cpp:58-80) ----
            GpuRhoSolver thisSolver(&mesh, &fLog, &allData);
            thisSolver.setDT(1. / STEP_TIME);
            thisSolver.solve();
            VCTDIMU* oldValue = thisSolver.getOldValue();
            VCTDIMU* newValue = thisSolver.getNewValue();
            VCTDIMU residual = VCTDIMU::Zero();
            for (int i = 0; i < nc; i++)
                for (int iR = 0; iR < DIMU; iR++)
                    if (residual[iR] < abs(newValue[i][iR] - oldValue[i][iR]) / oldValue[i][iR])
                        residual[iR] = abs(newValue[i][iR] - oldValue[i][iR]) / oldValue[i][iR];
            double dev[DIMU];
            thisSolver.residual(dev);  // same quantity, reduced on the device
            for (int k = 0; k < DIMU; k++) { resid.push_back(residual[k]); resid.push_back(dev[k]); }
            thisSolver.updateNewToOld();
        }
        char nm[32];
        snprintf(nm, sizeof nm, "Q%d", s);
        dump(nm, "f8", allData.getP1OldCellQs(), (long long)nc * (DIMU), 8);
    }
    dump("resid_host_dev", "f8", resid.data(), (long long)resid.size(), 8);
    {
        // output path: node-averaged fields of the final state from the device (GpuRhoSolverT::nodeFields),
        // fed by the reference's own Node / Face lists
        GpuRhoSolver outSolver(&mesh, &fLog, &allData);
        std::vector<double> nodeF;
        outSolver.nodeFields(nodeF);
        dump("node_fields", "f8", nodeF.data(), (long long)nodeF.size(), 8);
    }
    GpuRhoSolver::release(&mesh, &allData);
    fclose(g_out);
    return 0;
}
