// ORACLE / TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// Runs the REFERENCE'S OWN mesh reader (MshBlock::readMsh, R/mesh/MshBlock.cpp:5-16) and its own
// Tecplot writer (Work::writedataRhoBasedMshNodePlt, R/work/Work.cpp:204-319) -- both compiled from
// /root/reference through the symlink tree -- with wall-clock timing, so that the native reader
// (mst-cfd_b200/host/mshread.cpp) and the device node averaging + writer (csrc/mstgpu.cu k_node_fields,
// host/pltwrite.cpp) can be pinned to the reference's output byte for byte and timed beside it.
// Work's members are private and Work::work() never returns (stdin prompt): the driver opens the class
// with the usual test trick and calls the two members directly.
//
//   ref_io <mesh.msh> <outdir> <state.bin|-> <t> [<steps> [<flagmode>]]
//     outdir/result/ must exist; the writer names its file
//     outdir/result/<name>_TIME4000_u<inletu>_t<t>.plt (Work.cpp:221); the path is printed.
//     state: raw doubles ncells*DIMU, "-" = the SOD initial state of Time::initialization.
//     steps > 0: that many Time::goNextTimeStep calls (the reference's CPU RhoSolver: Roe, ACCURACY 2 in this
//     build) run before the file is written; the residual lines go to outdir/result/ref-log.lhblog
//     (Time.cpp:78).  flagmode 1 = `consistent` left/right flags (as in ref_driver.cpp), 0 = as the reader builds them.
// every standard / Eigen header the reference pulls in goes first, so the access trick below
// only touches the reference's own classes
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>
#include <Eigen/Eigen>
#include <Eigen/Sparse>
#include <Eigen/Dense>
#define private public
#include "work/Work.h"
#undef private

#include <chrono>
#include <cstdio>
#include <cstdlib>

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: ref_io mesh.msh outdir state.bin|- t\n"); return 2; }
    std::string msh = argv[1], outdir = argv[2], state = argv[3];
    const int tnow = atoi(argv[4]);
    const int nsteps = argc > 5 ? atoi(argv[5]) : 0, flagmode = argc > 6 ? atoi(argv[6]) : 0;
    Work w;
    w.mshAddress = outdir + "/";
    w.mshName = msh.substr(msh.find_last_of('/') + 1);
    FILE* keep = stdout;
    (void)keep;
    std::cout.setstate(std::ios_base::failbit);  // the reader's progress chatter
    double t0 = now_ms();
    w.mesh.readMsh(msh);
    const double read_ms = now_ms() - t0;
    const int nc = w.mesh.getNumOfCells(), nf = w.mesh.getNumOfFaces();
    w.allData.createAllData(nc, nf);
    VCTDIMU* Q = w.allData.getP1OldCellQs();
    if (flagmode == 1) {
        Face* faces = w.mesh.getBeginItFacesList();
        for (int f = 0; f < nf; f++)
            for (int d = 0; d < DIM; d++)
                faces[f].setFlagLeftRight(d, faces[f].getDirectAndCells() * faces[f].getDirect()[d] >= 0);
    }
    w.fLog.open(outdir + "/result/ref-log.lhblog", std::ios::out | std::ios::trunc);
    Time time1(&w.mesh, &w.fLog, &w.allData);
    {
        VCTDIMU iniQ;
        iniQ << inirho, inirho * iniu, inirho * iniv, iniE;
        time1.initialization(iniQ);
    }
    if (state != "-") {
        FILE* f = fopen(state.c_str(), "rb");
        if (!f || fread((void*)Q, sizeof(double) * (DIMU), (size_t)nc, f)  /* DIMU is "DIM + 2" unparenthesised, CONST.h:4 */ != (size_t)nc) { fprintf(stderr, "bad state file\n"); return 3; }
        fclose(f);
        for (int c = 0; c < nc; c++) w.allData.getP1NewCellQs()[c] = Q[c];
    }
    t0 = now_ms();
    for (int s = 0; s < nsteps; s++) time1.goNextTimeStep();
    const double step_ms = now_ms() - t0;
    w.fLog.flush();
    w.t = tnow;
    t0 = now_ms();
    w.writedataRhoBasedMshNodePlt(&w.mesh, Q, tnow);
    const double write_ms = now_ms() - t0;
    std::cout.clear();
    printf("cells %d faces %d nodes %d read_ms %.3f write_ms %.3f steps %d step_ms %.3f threads %d\n", nc, nf,
           w.mesh.getNumOfNodes(), read_ms, write_ms, nsteps, step_ms, NUM_CPU_THREADS);
    return 0;
}
