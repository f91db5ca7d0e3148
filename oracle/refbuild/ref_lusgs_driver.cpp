// ORACLE / TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// Runs the REFERENCE'S OWN LU-SGS solvers on a system read from a file:
//   block 1: SparseSolverNUM (R/lusolver/SparseSolverNUM.{h,cpp}) via setELE /
//            setRHSb / solveILUSGS / getPNewX
//   block 4: SparseSolver<MTDIMU_DIMU, VCTDIMU> (R/lusolver/SparseSolver.{h,cpp})
//            via setD / setL / setU / setRHSb / solveILU / getItBeginX
// Input file: int32 n, block, nnz; rowptr[n+1]; col[nnz]; val[nnz*block^2]
// (blocks row-major); b[n*block]; x0[n*block].  Output: x[n*block] raw doubles.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "time/Time.h"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* fi = fopen(argv[1], "rb");
    int hdr[3];
    if (!fi || fread(hdr, 4, 3, fi) != 3) return 3;
    const int n = hdr[0], B = hdr[1], nnz = hdr[2];
    std::vector<int> rowptr(n + 1), col(nnz);
    std::vector<double> val((size_t)nnz * B * B), b((size_t)n * B), x0((size_t)n * B);
    if (fread(rowptr.data(), 4, n + 1, fi) != (size_t)n + 1 || fread(col.data(), 4, nnz, fi) != (size_t)nnz ||
        fread(val.data(), 8, val.size(), fi) != val.size() || fread(b.data(), 8, b.size(), fi) != b.size() ||
        fread(x0.data(), 8, x0.size(), fi) != x0.size())
        return 4;
    fclose(fi);
    std::vector<double> x((size_t)n * B);
    if (B == 1) {
        SparseSolverNUM s(n, x0.data());
        for (int r = 0; r < n; r++)
            for (int k = rowptr[r]; k < rowptr[r + 1]; k++) s.setELE(val[k], r, col[k]);
        for (int r = 0; r < n; r++) s.setRHSb(b[r], r);
        s.solveILUSGS();
        for (int r = 0; r < n; r++) x[r] = s.getPNewX()[r];
    } else if (B == DIMU) {
        std::vector<VCTDIMU> old(n);
        for (int r = 0; r < n; r++)
            for (int q = 0; q < B; q++) old[r][q] = x0[(size_t)r * B + q];
        SparseSolver<MTDIMU_DIMU, VCTDIMU> s(n, old.data());
        for (int r = 0; r < n; r++)
            for (int k = rowptr[r]; k < rowptr[r + 1]; k++) {
                MTDIMU_DIMU m;
                for (int i = 0; i < B; i++)
                    for (int j = 0; j < B; j++) m(i, j) = val[(size_t)k * B * B + i * B + j];
                if (col[k] == r) s.setD(m, r);
                else if (col[k] < r) s.setL(m, r, col[k]);
                else s.setU(m, r, col[k]);
            }
        for (int r = 0; r < n; r++) {
            VCTDIMU v;
            for (int q = 0; q < B; q++) v[q] = b[(size_t)r * B + q];
            s.setRHSb(v, r);
        }
        s.solveILU();
        for (int r = 0; r < n; r++)
            for (int q = 0; q < B; q++) x[(size_t)r * B + q] = s.getItBeginX()[r][q];
    } else {
        return 5;
    }
    FILE* fo = fopen(argv[2], "wb");
    fwrite(x.data(), 8, x.size(), fo);
    fclose(fo);
    return 0;
}
