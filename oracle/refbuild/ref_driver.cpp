// ORACLE / TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// Driver that runs the REFERENCE'S OWN rhoSolver path (its MshBlock reader,
// AllData, Time::goNextTimeStep, RhoSolver, solverRoe / SolverAusm -- compiled
// from /root/reference through the symlink tree the Makefile builds) and dumps
// what the oracle restatement has to reproduce.  It replaces R/main.cpp and
// R/work/Work.cpp (hard-coded mesh name, Tecplot output, stdin prompt) only.
//
//   ref_<variant> <mesh.msh> <out.bin> <flagmode> <retag> <init.bin|-> <step> [<step> ...]
//     flagmode 0: flags as the reader builds them (MshBlock.cpp:281-305)
//              1: `consistent` flags, set through Face::setFlagLeftRight
//     retag    "a:b" rewrites zone type a to b through FacesInf::setType ("-": none)
//     init     raw doubles ncells*DIMU replacing the SOD initial state ("-": keep)
//     steps    ascending step counts at which Q is dumped
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "time/Time.h"

static FILE* g_out;
static void dump(const char* name, const char* dtype, const void* p, long long count, int width) {
    fprintf(g_out, "%s %s %lld\n", name, dtype, count);
    fwrite(p, width, (size_t)count, g_out);
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: see source\n"); return 2; }
    std::string msh = argv[1];
    g_out = fopen(argv[2], "wb");
    int flagmode = atoi(argv[3]);
    std::string retag = argv[4], init = argv[5];
    std::vector<int> steps;
    for (int i = 6; i < argc; i++) steps.push_back(atoi(argv[i]));

    MshBlock mesh;
    mesh.readMsh(msh);
    const int nc = mesh.getNumOfCells(), nf = mesh.getNumOfFaces(), nint = mesh.getNumOfIntFaces();
    Face* faces = mesh.getBeginItFacesList();
    Cell* cells = mesh.getBeginItCellsList();
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

    // ---- mesh tables as the solver reads them -------------------------------
    {
        std::vector<double> S(nf * DIM), fc(nf * DIM), eta(nf), cc(nc * DIM), vol(nc), sout;
        std::vector<int> c0(nf), c1(nf), ftype(nf, 0), cfptr(nc + 1), cfidx;
        std::vector<signed char> dac(nf);
        std::vector<unsigned char> flag(nf * DIM);
        for (int f = 0; f < nf; f++) {
            for (int d = 0; d < DIM; d++) {
                S[f * DIM + d] = faces[f].getDirect()[d];
                fc[f * DIM + d] = faces[f].getCenter()[d];
                flag[f * DIM + d] = faces[f].getFlagLeftRight()[d] ? 1 : 0;
            }
            eta[f] = faces[f].getEta0();
            dac[f] = (signed char)faces[f].getDirectAndCells();
            c0[f] = faces[f].getBeginItPNbCells()[0]->getId();
            c1[f] = faces[f].getNumOfpNbCells() == 2 ? faces[f].getBeginItPNbCells()[1]->getId() : -1;
        }
        auto it = mesh.getBeginItFacesInfList();
        for (int z = 0; z < mesh.getNumOfFacesInfs(); z++)
            for (int f = it[z].getStart() - 1; f < it[z].getEnd(); f++) ftype[f] = it[z].getType();
        cfptr[0] = 0;
        for (int c = 0; c < nc; c++) {
            vol[c] = cells[c].getVolume();
            for (int d = 0; d < DIM; d++) cc[c * DIM + d] = cells[c].getCenter()[d];
            for (int j = 0; j < cells[c].getNumOfNbFaces(); j++) {
                cfidx.push_back(cells[c].getBeginItPNbFaces()[j]->getId());
                for (int d = 0; d < DIM; d++) sout.push_back(cells[c].getBeginItDirectOfNbFaces()[j][d]);
            }
            cfptr[c + 1] = (int)cfidx.size();
        }
        int hdr[4] = {DIM, nc, nf, nint};
        dump("hdr", "i4", hdr, 4, 4);
        dump("c0", "i4", c0.data(), nf, 4); dump("c1", "i4", c1.data(), nf, 4);
        dump("S", "f8", S.data(), nf * DIM, 8); dump("fc", "f8", fc.data(), nf * DIM, 8);
        dump("eta", "f8", eta.data(), nf, 8); dump("dac", "i1", dac.data(), nf, 1);
        dump("flag", "u1", flag.data(), nf * DIM, 1); dump("ftype", "i4", ftype.data(), nf, 4);
        dump("cc", "f8", cc.data(), nc * DIM, 8); dump("vol", "f8", vol.data(), nc, 8);
        dump("cf_ptr", "i4", cfptr.data(), nc + 1, 4); dump("cf_idx", "i4", cfidx.data(), (long long)cfidx.size(), 4);
        dump("sout", "f8", sout.data(), (long long)sout.size(), 8);
    }

    AllData allData;
    allData.createAllData(nc, nf);
    std::fstream fLog("/dev/null", std::ios::out);
    Time time1(&mesh, &fLog, &allData);
    VCTDIMU iniQ;
    iniQ << inirho, inirho * iniu, inirho * iniv, iniE;  // R/work/Work.cpp:34
    time1.initialization(iniQ);
    if (init != "-") {
        std::vector<double> q((size_t)nc * (DIMU));
        FILE* fi = fopen(init.c_str(), "rb");
        if (!fi || fread(q.data(), 8, q.size(), fi) != q.size()) { fprintf(stderr, "bad init file\n"); return 3; }
        fclose(fi);
        for (int c = 0; c < nc; c++)
            for (int k = 0; k < DIMU; k++) {
                allData.getP1OldCellQs()[c][k] = q[(size_t)c * (DIMU) + k];
                allData.getP1NewCellQs()[c][k] = q[(size_t)c * (DIMU) + k];
            }
    }
    dump("Q0", "f8", allData.getP1OldCellQs(), (long long)nc * (DIMU), 8);
    int done = 0;
    for (int s : steps) {
        for (; done < s; done++) time1.goNextTimeStep();
        char nm[32];
        snprintf(nm, sizeof nm, "Q%d", s);
        dump(nm, "f8", allData.getP1OldCellQs(), (long long)nc * (DIMU), 8);
        if (s == steps[0]) {
            // face flux matrices (DIMU x DIM, column-major) and face values of the last solve
            dump("F", "f8", allData.getP1OldFaceConvectFlux(), (long long)nf * (DIMU) * DIM, 8);
            dump("Qf", "f8", allData.getP1OldFaceQs(), (long long)nf * (DIMU), 8);
        }
    }
    fclose(g_out);
    return 0;
}
