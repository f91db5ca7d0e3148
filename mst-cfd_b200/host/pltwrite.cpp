// pltwrite.cpp -- host half of the output path.  R = /root/reference/MST-CFD.
//
// The reference writes a node-averaged ASCII Tecplot file every 10 steps
// (Work::writedataRhoBasedMshNodePlt, R/work/Work.cpp:204-319): face states from the cell states,
// node states from the faces around each node, primitives per node, then the element list, every
// number through `ostream << fixed << setprecision(15)`.  Here the arithmetic runs on the device
// (csrc/output.cuh, mstgpu_node_fields: one thread per node, no face array, only the node fields
// come back over PCIe) and this file turns the result into the SAME BYTES the reference writes:
//
//   msthost_cell_nodes   Cell::getBeginItPNbNodes as MshBlock.cpp:335-368 builds it: nodes in
//                        first-seen order over the cell's faces (file order) and each face's nodes;
//                        a 4-node cell swaps its last two nodes when (n0-n1).(n2-n3) < 0
//   msthost_plt_write    the file itself, multi-threaded formatting (to_chars, fixed, 15 digits ==
//                        what num_put/printf("%.15f") produces: both are exact decimal expansions)
//   msthost_plt_write_binary  the same content as raw little-endian doubles / int32 (what 8f.3 asks
//                        for when nobody needs the 15-digit text)
//
// Pinned byte for byte against the reference's own writer compiled from /root/reference
// (oracle/refbuild/ref_io_driver.cpp; tests/test_output_cpu.py, tests/test_output_gpu.py).
#include <omp.h>

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "textout.h"

namespace {

// ostream << fixed << setprecision(15) [<< setw(15)] << v.  A finite number has 17 characters at least, so
// setw(15) only shows on nan / inf, which are right-aligned in 15 columns like iostream does.
inline char* put_fixed15(char* p, double v, bool setw15) {
    if (std::isfinite(v)) return std::to_chars(p, p + 340, v, std::chars_format::fixed, 15).ptr;
    char t[16];
    const int n = snprintf(t, sizeof t, "%.15f", v);  // nan / -nan / inf / -inf exactly as printf spells them
    if (setw15) for (int k = n; k < 15; k++) *p++ = ' ';
    memcpy(p, t, (size_t)n);
    return p + n;
}

}  // namespace

extern "C" {

// cn_ptr [ncells+1] is always filled; cn_idx (capacity cn_ptr[ncells]) may be null on a sizing call.
int msthost_cell_nodes(int32_t dim, int64_t ncells, int32_t npf, const int32_t* face_nodes, const int32_t* cf_ptr,
                       const int32_t* cf_idx, const double* nodes, int32_t* cn_ptr, int32_t* cn_idx) {
    if (!face_nodes || !cf_ptr || !cf_idx || !cn_ptr || (cn_idx && !nodes)) return -1;
    const int D = dim;
    std::vector<int32_t> cnt((size_t)ncells);
    auto gather = [&](int64_t c, int32_t* out) {  // returns the number of distinct nodes
        int n = 0;
        for (int j = cf_ptr[c]; j < cf_ptr[c + 1]; j++) {
            const int32_t* fn = face_nodes + (int64_t)cf_idx[j] * npf;
            for (int k = 0; k < npf; k++) {
                const int32_t v = fn[k];
                if (v < 0) continue;
                bool have = false;
                for (int q = 0; q < n; q++) have |= out[q] == v;
                if (!have && n < 16) out[n++] = v;
            }
        }
        return n;
    };
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncells; c++) {
        int32_t tmp[16];
        cnt[(size_t)c] = gather(c, tmp);
    }
    cn_ptr[0] = 0;
    for (int64_t c = 0; c < ncells; c++) cn_ptr[c + 1] = cn_ptr[c] + cnt[(size_t)c];
    if (!cn_idx) return 0;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncells; c++) {
        int32_t tmp[16];
        const int n = gather(c, tmp);
        if (n == 4) {  // MshBlock.cpp:358-367
            double dot = 0.0;
            for (int d = 0; d < D; d++) {
                const double a = nodes[(int64_t)tmp[0] * D + d] - nodes[(int64_t)tmp[1] * D + d];
                const double b = nodes[(int64_t)tmp[2] * D + d] - nodes[(int64_t)tmp[3] * D + d];
                dot = d ? dot + a * b : a * b;
            }
            if (dot < 0) std::swap(tmp[2], tmp[3]);
        }
        memcpy(cn_idx + cn_ptr[c], tmp, sizeof(int32_t) * (size_t)n);
    }
    return 0;
}

// fields [nnodes][dim+4] = rho, u_i, T, p, Ma (what mstgpu_node_fields returns).  zone_t is the
// step counter the reference prints in ZONE T="..." (Work::t); felnum = FELNUM of CONST.h:5
// (3 -> FETRIANGLE, anything else -> FEQUADRILATERAL; 3-D always FETETRAHEDRON).
int msthost_plt_write(const char* path, int32_t dim, int64_t nnodes, int64_t ncells, const double* nodes,
                      const double* fields, const int32_t* cn_ptr, const int32_t* cn_idx, int32_t zone_t,
                      int32_t felnum) {
    if (!path || !nodes || !fields || !cn_ptr || !cn_idx || (dim != 2 && dim != 3)) return -1;
    FILE* fp = fopen(path, "wb");
    if (!fp) return -2;
    const int D = dim, W = dim + 4;
    fputs("\"TITLE = \"Example: Variable and Connectivity List Sharing\"\n", fp);
    if (D == 2) {
        fputs("VARIABLES = \"X\", \"Y\", \"rho\" , \"u\" , \"v\" , \"T\" , \"p\" , \"Ma\" \n", fp);
        fprintf(fp, "ZONE T=\"%d\", DATAPACKING=POINT, NODES=%lld, ELEMENTS=%lld, ZONETYPE=%s\n", zone_t,
                (long long)nnodes, (long long)ncells, felnum == 3 ? "FETRIANGLE" : "FEQUADRILATERAL");
    } else {
        fputs("VARIABLES = \"X\", \"Y\", \"Z\", \"rho\" , \"u\" , \"v\" , \"T\" , \"p\" , \"Ma\" \n", fp);
        fprintf(fp, "ZONE T=\"%d\", DATAPACKING=POINT, NODES=%lld, ELEMENTS=%lld, ZONETYPE=FETETRAHEDRON\n", zone_t,
                (long long)nnodes, (long long)ncells);
    }
    // node lines (Work.cpp:296-304): coordinates, rho, u_i, T, p, Ma, each followed by one blank
    // widest number in the file: sign + integer digits + '.' + 15 decimals
    double big = 1.0;
#pragma omp parallel for schedule(static) reduction(max : big)
    for (int64_t i = 0; i < nnodes; i++) {
        for (int d = 0; d < D; d++) { const double a = std::fabs(nodes[i * D + d]); if (std::isfinite(a) && a > big) big = a; }
        for (int k = 0; k < W; k++) { const double a = std::fabs(fields[i * W + k]); if (std::isfinite(a) && a > big) big = a; }
    }
    const int numw = (int)std::floor(std::log10(big)) + 1 + 1 + 1 + 15 + 2;
    bool ok = msthost::emit_records(fp, nnodes, (D + W) * (numw + 1) + 2, [&](int64_t i, char* p) {
        // setw(15) is in front of the coordinates, rho, u_i and T only (Work.cpp:297-303)
        for (int d = 0; d < D; d++) { p = put_fixed15(p, nodes[i * D + d], true); *p++ = ' '; }
        for (int k = 0; k < W; k++) { p = put_fixed15(p, fields[i * W + k], k < D + 2); *p++ = ' '; }
        *p++ = '\n';
        return p;
    });
    // element lines (Work.cpp:306-312): 1-based node ids, each followed by one blank
    int maxn = 0;
    for (int64_t c = 0; c < ncells; c++) maxn = std::max(maxn, cn_ptr[c + 1] - cn_ptr[c]);
    ok = ok && msthost::emit_records(fp, ncells, 11 * std::max(maxn, 1) + 2, [&](int64_t c, char* p) {
        for (int j = cn_ptr[c]; j < cn_ptr[c + 1]; j++) { p = msthost::put_dec(p, (uint32_t)(cn_idx[j] + 1)); *p++ = ' '; }
        *p++ = '\n';
        return p;
    });
    ok = (fclose(fp) == 0) && ok;
    return ok ? 0 : -2;
}

// Binary twin: "MSTPLT1\0", int32 dim, int32 zone_t, int64 nnodes, int64 ncells, int64 nconn, then
// nodes [nnodes*dim] f64, fields [nnodes*(dim+4)] f64, cn_ptr [ncells+1] i32, cn_idx [nconn] i32.
int msthost_plt_write_binary(const char* path, int32_t dim, int64_t nnodes, int64_t ncells, const double* nodes,
                             const double* fields, const int32_t* cn_ptr, const int32_t* cn_idx, int32_t zone_t) {
    if (!path || !nodes || !fields || !cn_ptr || !cn_idx || (dim != 2 && dim != 3)) return -1;
    FILE* fp = fopen(path, "wb");
    if (!fp) return -2;
    const int64_t nconn = cn_ptr[ncells];
    bool ok = fwrite("MSTPLT1", 1, 8, fp) == 8;
    ok = ok && fwrite(&dim, 4, 1, fp) == 1 && fwrite(&zone_t, 4, 1, fp) == 1;
    ok = ok && fwrite(&nnodes, 8, 1, fp) == 1 && fwrite(&ncells, 8, 1, fp) == 1 && fwrite(&nconn, 8, 1, fp) == 1;
    ok = ok && fwrite(nodes, 8, (size_t)(nnodes * dim), fp) == (size_t)(nnodes * dim);
    ok = ok && fwrite(fields, 8, (size_t)(nnodes * (dim + 4)), fp) == (size_t)(nnodes * (dim + 4));
    ok = ok && fwrite(cn_ptr, 4, (size_t)(ncells + 1), fp) == (size_t)(ncells + 1);
    ok = ok && fwrite(cn_idx, 4, (size_t)nconn, fp) == (size_t)nconn;
    ok = (fclose(fp) == 0) && ok;
    return ok ? 0 : -2;
}

}  // extern "C"
