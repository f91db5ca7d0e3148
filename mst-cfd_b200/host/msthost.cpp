// msthost.cpp -- host-side companions of the GPU path (C ABI, no CUDA):
//
//  * msthost_flatten_*   raw mesh tables (nodes, face->nodes, c0, c1, zone type)
//                        -> the flat tables of include/mstgpu.h (mstgpu_mesh),
//                        i.e. what the reference's MshBlock computes after
//                        reading a file (R = /root/reference/MST-CFD):
//                        R/mesh/Face.cpp:8-44,62-69, R/mesh/Cell.cpp:6-61,
//                        R/mesh/MshBlock.cpp:281-334.  In a drop-in build the
//                        reference's own MshBlock supplies these numbers; this
//                        flattener is for hosts that only have raw tables
//                        (synthetic meshes, the bench) and is multi-threaded so
//                        that a 50 M-cell mesh is ready in seconds.
//  * msthost_box_tets    synthetic n_x * n_y * n_z hex box, Kuhn 6-tet split
//                        (BASELINE configs 4-5: 203^3 -> 50 192 562 tets).
//  * msthost_grid_tris   synthetic structured 2-D grid with a cell mask, each
//                        quad split along a fixed diagonal (BASELINE config 2:
//                        forward-facing step, 998 046 triangles).
//
// 3-D metrics are the documented extension: V = 1/3 sum Sout.(fc-cc),
// flags by the `consistent` rule.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

// ---------------------------------------------------------------------------
// sizes of the Kuhn-split box
void msthost_box_tets_sizes(int nx, int ny, int nz, int64_t* nnodes, int64_t* ncells,
                            int64_t* nfaces, int64_t* nint) {
    const int64_t X = nx, Y = ny, Z = nz;
    *nnodes = (X + 1) * (Y + 1) * (Z + 1);
    *ncells = 6 * X * Y * Z;
    const int64_t inner = 6 * X * Y * Z + 2 * ((X - 1) * Y * Z + X * (Y - 1) * Z + X * Y * (Z - 1));
    const int64_t bnd = 4 * (Y * Z + X * Z + X * Y);
    *nint = inner;
    *nfaces = inner + bnd;
}

// perms of the axes, tet t of a hex has vertices 0, e_p0, e_p0+e_p1, (1,1,1)
static const int PERM[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
static int perm_index(int a, int b, int c) {
    for (int t = 0; t < 6; t++)
        if (PERM[t][0] == a && PERM[t][1] == b && PERM[t][2] == c) return t;
    return -1;
}

// bc[6] = zone types of the sides x-, x+, y-, y+, z-, z+.
// Output: nodes [nnodes*3], face_nodes [nfaces*3], c0, c1 [nfaces] (c1 = -1 on
// the boundary), ftype [nfaces].  Interior faces come first (hex-major), then
// the boundary faces side by side, like a Fluent file.
int msthost_box_tets(int nx, int ny, int nz, double lx, double ly, double lz, const int32_t* bc,
                     double* nodes, int32_t* face_nodes, int32_t* c0, int32_t* c1, int32_t* ftype) {
    const int64_t NX1 = nx + 1, NY1 = ny + 1;
    const int n[3] = {nx, ny, nz};
    auto node_id = [&](int i, int j, int k) -> int32_t { return (int32_t)(((int64_t)k * NY1 + j) * NX1 + i); };
    auto hex_id = [&](int i, int j, int k) -> int64_t { return ((int64_t)k * ny + j) * nx + i; };
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz; k++)
        for (int j = 0; j <= ny; j++)
            for (int i = 0; i <= nx; i++) {
                const int64_t id = node_id(i, j, k);
                nodes[id * 3 + 0] = lx * i / nx;
                nodes[id * 3 + 1] = ly * j / ny;
                nodes[id * 3 + 2] = lz * k / nz;
            }
    // interior faces: per hex 6 internal + 2 per existing low-side neighbour
    const int64_t nhex = (int64_t)nx * ny * nz;
    std::vector<int64_t> off(nhex + 1);
    off[0] = 0;
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const int64_t h = hex_id(i, j, k);
                off[h + 1] = off[h] + 6 + 2 * ((i > 0) + (j > 0) + (k > 0));
            }
    const int64_t nint = off[nhex];
    // internal pairs: "remove p1" pairs perms that differ by swapping the first
    // two axes, "remove p2" pairs perms that differ by swapping the last two.
#pragma omp parallel for schedule(static) collapse(2)
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const int64_t h = hex_id(i, j, k);
                int64_t f = off[h];
                const int base[3] = {i, j, k};
                auto vtx = [&](const int* a) { return node_id(base[0] + a[0], base[1] + a[1], base[2] + a[2]); };
                for (int t = 0; t < 6; t++) {
                    const int* p = PERM[t];
                    int v0[3] = {0, 0, 0}, v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0}, v3[3] = {1, 1, 1};
                    v1[p[0]] = 1;
                    v2[p[0]] = 1; v2[p[1]] = 1;
                    // remove p1: {p0,p2,p3}, partner swaps first two axes
                    int u = perm_index(p[1], p[0], p[2]);
                    if (t < u) {
                        face_nodes[f * 3 + 0] = vtx(v0); face_nodes[f * 3 + 1] = vtx(v2); face_nodes[f * 3 + 2] = vtx(v3);
                        c0[f] = (int32_t)(6 * h + t); c1[f] = (int32_t)(6 * h + u); ftype[f] = 2; f++;
                    }
                    // remove p2: {p0,p1,p3}, partner swaps last two axes
                    u = perm_index(p[0], p[2], p[1]);
                    if (t < u) {
                        face_nodes[f * 3 + 0] = vtx(v0); face_nodes[f * 3 + 1] = vtx(v1); face_nodes[f * 3 + 2] = vtx(v3);
                        c0[f] = (int32_t)(6 * h + t); c1[f] = (int32_t)(6 * h + u); ftype[f] = 2; f++;
                    }
                }
                // low-side faces shared with the neighbour hex at -e_x: this
                // hex's tet (a,b,x) <-> neighbour's tet (x,a,b)
                for (int x = 0; x < 3; x++) {
                    if (base[x] == 0) continue;
                    int nb[3] = {i, j, k};
                    nb[x] -= 1;
                    const int64_t hn = hex_id(nb[0], nb[1], nb[2]);
                    const int a = (x + 1) % 3, b = (x + 2) % 3;
                    const int pr[2][2] = {{a, b}, {b, a}};
                    for (int q = 0; q < 2; q++) {
                        int v0[3] = {0, 0, 0}, v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0};
                        v1[pr[q][0]] = 1;
                        v2[pr[q][0]] = 1; v2[pr[q][1]] = 1;
                        face_nodes[f * 3 + 0] = vtx(v0); face_nodes[f * 3 + 1] = vtx(v1); face_nodes[f * 3 + 2] = vtx(v2);
                        c0[f] = (int32_t)(6 * hn + perm_index(x, pr[q][0], pr[q][1]));
                        c1[f] = (int32_t)(6 * h + perm_index(pr[q][0], pr[q][1], x));
                        ftype[f] = 2;
                        f++;
                    }
                }
            }
    // boundary faces, side by side
    int64_t f = nint;
    for (int x = 0; x < 3; x++) {
        const int a = (x + 1) % 3, b = (x + 2) % 3;
        for (int side = 0; side < 2; side++) {
            const int64_t cnt = 2 * (int64_t)n[a] * n[b];
            const int64_t f0 = f;
#pragma omp parallel for schedule(static)
            for (int64_t q = 0; q < (int64_t)n[a] * n[b]; q++) {
                int idx[3];
                idx[a] = (int)(q % n[a]);
                idx[b] = (int)(q / n[a]);
                idx[x] = side ? n[x] - 1 : 0;
                const int64_t h = hex_id(idx[0], idx[1], idx[2]);
                const int pr[2][2] = {{a, b}, {b, a}};
                for (int s = 0; s < 2; s++) {
                    const int64_t ff = f0 + 2 * q + s;
                    int v0[3] = {0, 0, 0}, v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0};
                    int t;
                    if (side == 0) {  // low side: tet (a,b,x), face {0, e_a, e_a+e_b}
                        v1[pr[s][0]] = 1;
                        v2[pr[s][0]] = 1; v2[pr[s][1]] = 1;
                        t = perm_index(pr[s][0], pr[s][1], x);
                    } else {  // high side: tet (x,a,b), face {e_x, e_x+e_a, 111}
                        v0[x] = 1;
                        v1[x] = 1; v1[pr[s][0]] = 1;
                        v2[0] = v2[1] = v2[2] = 1;
                        t = perm_index(x, pr[s][0], pr[s][1]);
                    }
                    face_nodes[ff * 3 + 0] = node_id(idx[0] + v0[0], idx[1] + v0[1], idx[2] + v0[2]);
                    face_nodes[ff * 3 + 1] = node_id(idx[0] + v1[0], idx[1] + v1[1], idx[2] + v1[2]);
                    face_nodes[ff * 3 + 2] = node_id(idx[0] + v2[0], idx[1] + v2[1], idx[2] + v2[2]);
                    c0[ff] = (int32_t)(6 * h + t);
                    c1[ff] = -1;
                    ftype[ff] = bc[2 * x + side];
                }
            }
            f += cnt;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// 2-D structured grid nx*ny over [0,lx]x[0,ly]; mask[j*nx+i] != 0 marks active
// quads; each active quad is split into two triangles along the (i,j)-(i+1,j+1)
// diagonal.  bc[3] = {type at x = 0, type at x = lx, type elsewhere}.
void msthost_grid_tris_sizes(int nx, int ny, const uint8_t* mask, int64_t* nnodes, int64_t* ncells,
                             int64_t* nfaces, int64_t* nint) {
    int64_t nq = 0, ni = 0, nb = 0;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            if (!mask[(int64_t)j * nx + i]) continue;
            nq++;
            ni++;  // the diagonal
            // edges: left and bottom owned by this quad; right/top when the neighbour is missing
            const bool l = i > 0 && mask[(int64_t)j * nx + i - 1];
            const bool b = j > 0 && mask[(int64_t)(j - 1) * nx + i];
            const bool r = i + 1 < nx && mask[(int64_t)j * nx + i + 1];
            const bool t = j + 1 < ny && mask[(int64_t)(j + 1) * nx + i];
            (l ? ni : nb)++;
            (b ? ni : nb)++;
            if (!r) nb++;
            if (!t) nb++;
        }
    *nnodes = (int64_t)(nx + 1) * (ny + 1);
    *ncells = 2 * nq;
    *nint = ni;
    *nfaces = ni + nb;
}

int msthost_grid_tris(int nx, int ny, double lx, double ly, const uint8_t* mask, const int32_t* bc,
                      double* nodes, int32_t* face_nodes, int32_t* c0, int32_t* c1, int32_t* ftype) {
    const int64_t NX1 = nx + 1;
    auto nid = [&](int i, int j) -> int32_t { return (int32_t)((int64_t)j * NX1 + i); };
    for (int j = 0; j <= ny; j++)
        for (int i = 0; i <= nx; i++) {
            nodes[2 * (int64_t)nid(i, j) + 0] = lx * i / nx;
            nodes[2 * (int64_t)nid(i, j) + 1] = ly * j / ny;
        }
    std::vector<int32_t> qid((int64_t)nx * ny, -1);
    int32_t nq = 0;
    for (int64_t q = 0; q < (int64_t)nx * ny; q++)
        if (mask[q]) qid[q] = nq++;
    // triangle 0 of a quad = lower-right (nodes (i,j),(i+1,j),(i+1,j+1)): owns bottom + right edges
    // triangle 1          = upper-left  (nodes (i,j),(i+1,j+1),(i,j+1)): owns left + top edges
    int64_t nint, nfaces, nn, ncell;
    msthost_grid_tris_sizes(nx, ny, mask, &nn, &ncell, &nfaces, &nint);
    int64_t fi = 0, fb = nint;
    auto put = [&](int64_t f, int32_t a, int32_t b, int32_t ca, int32_t cb, int32_t ty) {
        face_nodes[2 * f] = a; face_nodes[2 * f + 1] = b; c0[f] = ca; c1[f] = cb; ftype[f] = ty;
    };
    auto btype = [&](int i0, int i1) -> int32_t {
        if (i0 == 0 && i1 == 0) return bc[0];
        if (i0 == nx && i1 == nx) return bc[1];
        return bc[2];
    };
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) {
            const int32_t q = qid[(int64_t)j * nx + i];
            if (q < 0) continue;
            const int32_t t0 = 2 * q, t1 = 2 * q + 1;
            put(fi++, nid(i, j), nid(i + 1, j + 1), t0, t1, 2);  // diagonal
            const int32_t ql = i > 0 ? qid[(int64_t)j * nx + i - 1] : -1;
            const int32_t qb = j > 0 ? qid[(int64_t)(j - 1) * nx + i] : -1;
            const int32_t qr = i + 1 < nx ? qid[(int64_t)j * nx + i + 1] : -1;
            const int32_t qt = j + 1 < ny ? qid[(int64_t)(j + 1) * nx + i] : -1;
            if (ql >= 0) put(fi++, nid(i, j), nid(i, j + 1), 2 * ql, t1, 2);  // left nb's tri 0 owns its right edge
            else put(fb++, nid(i, j), nid(i, j + 1), t1, -1, btype(i, i));
            if (qb >= 0) put(fi++, nid(i, j), nid(i + 1, j), 2 * qb + 1, t0, 2);  // lower nb's tri 1 owns its top edge
            else put(fb++, nid(i, j), nid(i + 1, j), t0, -1, bc[2]);
            if (qr < 0) put(fb++, nid(i + 1, j), nid(i + 1, j + 1), t0, -1, btype(i + 1, i + 1));
            if (qt < 0) put(fb++, nid(i, j + 1), nid(i + 1, j + 1), t1, -1, bc[2]);
        }
    return (fi == nint && fb == nfaces) ? 0 : -1;
}

// ---------------------------------------------------------------------------
// flattener: raw tables -> flat mesh (reference order).  npf = nodes per face
// (row stride of face_nodes; entries < 0 = unused).  flag_convention: 0 =
// consistent (flag[d] = Sout_c0[d] >= 0), 1 = as shipped (MshBlock.cpp:284-303).
// Outputs (caller-allocated): S, fc [nf*dim], dac, eta [nf], flag [nf*dim],
// cc [nc*dim], vol [nc], cf_ptr [nc+1], cf_idx [sum faces per cell].
int msthost_flatten(int dim, int64_t nnodes, int64_t ncells, int64_t nfaces, int npf,
                    const double* nodes, const int32_t* face_nodes, const int32_t* c0,
                    const int32_t* c1, int flag_convention, double* S, double* fc, int8_t* dac,
                    double* eta, uint8_t* flag, double* cc, double* vol, int32_t* cf_ptr,
                    int32_t* cf_idx) {
    const int D = dim;
    if (D != 2 && D != 3) return -1;
    (void)nnodes;
    // face centre + area vector (Face.cpp:8-44)
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < nfaces; f++) {
        const int32_t* fn = face_nodes + f * npf;
        int cnt = 0;
        double s[3] = {0, 0, 0};
        for (int k = 0; k < npf; k++) {
            if (fn[k] < 0) continue;
            for (int d = 0; d < D; d++) s[d] += nodes[(int64_t)fn[k] * D + d];
            cnt++;
        }
        for (int d = 0; d < D; d++) fc[f * D + d] = s[d] / (double)cnt;
        if (D == 2) {
            const double fx = nodes[(int64_t)fn[0] * 2] - nodes[(int64_t)fn[1] * 2];
            const double fy = nodes[(int64_t)fn[0] * 2 + 1] - nodes[(int64_t)fn[1] * 2 + 1];
            S[f * 2] = -fy;
            S[f * 2 + 1] = fx;
        } else {
            double a[3], b[3];
            for (int d = 0; d < 3; d++) {
                a[d] = nodes[(int64_t)fn[1] * 3 + d] - nodes[(int64_t)fn[0] * 3 + d];
                b[d] = nodes[(int64_t)fn[2] * 3 + d] - nodes[(int64_t)fn[0] * 3 + d];
            }
            const double h = (cnt == 3) ? 0.5 : 1.0;  // quad: no 1/2 (Face.cpp:30-35)
            S[f * 3 + 0] = h * (a[1] * b[2] - a[2] * b[1]);
            S[f * 3 + 1] = h * (a[2] * b[0] - a[0] * b[2]);
            S[f * 3 + 2] = h * (a[0] * b[1] - a[1] * b[0]);
        }
    }
    // cell -> faces in file order (MshBlock.cpp:238-239,254-255)
    std::vector<int32_t> cnt(ncells + 1, 0);
    for (int64_t f = 0; f < nfaces; f++) {
        cnt[c0[f] + 1]++;
        if (c1[f] >= 0) cnt[c1[f] + 1]++;
    }
    cf_ptr[0] = 0;
    for (int64_t c = 0; c < ncells; c++) cf_ptr[c + 1] = cf_ptr[c] + cnt[c + 1];
    std::vector<int32_t> pos(cf_ptr, cf_ptr + ncells);
    for (int64_t f = 0; f < nfaces; f++) {
        cf_idx[pos[c0[f]]++] = (int32_t)f;
        if (c1[f] >= 0) cf_idx[pos[c1[f]]++] = (int32_t)f;
    }
    auto nrm = [&](const double* v) {
        double s = v[0] * v[0];
        for (int d = 1; d < D; d++) s = s + v[d] * v[d];
        return std::sqrt(s);
    };
    // cell centre = mean of face centres, summed in list order (Cell.cpp:6-14);
    // 2-D volume by Heron (Cell.cpp:15-51)
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncells; c++) {
        const int b = cf_ptr[c], e = cf_ptr[c + 1], nfc = e - b;
        double s[3] = {0, 0, 0};
        for (int j = b; j < e; j++)
            for (int d = 0; d < D; d++) s[d] += fc[(int64_t)cf_idx[j] * D + d];
        for (int d = 0; d < D; d++) cc[c * D + d] = s[d] / (double)nfc;
        if (D == 2) {
            double v = 0.0;
            if (nfc == 3) {
                double l[3];
                for (int k = 0; k < 3; k++) l[k] = nrm(S + (int64_t)cf_idx[b + k] * 2);
                const double h = 0.5 * (l[0] + l[1] + l[2]);
                v = std::sqrt(h * (h - l[0]) * (h - l[1]) * (h - l[2]));
            } else if (nfc == 4) {
                const int32_t* fl = cf_idx + b;
                auto fcx = [&](int k, int d) { return fc[(int64_t)fl[k] * 2 + d]; };
                const double v1[2] = {fcx(0, 0) - fcx(1, 0), fcx(0, 1) - fcx(1, 1)};
                const double v2[2] = {fcx(2, 0) - fcx(3, 0), fcx(2, 1) - fcx(3, 1)};
                double l[4];
                for (int k = 0; k < 4; k++) l[k] = nrm(S + (int64_t)fl[k] * 2);
                int p0, p1, p2, p3, other;
                if (std::fabs(v1[0] * v2[1] - v2[0] * v1[1]) < 1e-7) { p0 = 0; p1 = 1; p2 = 2; p3 = 3; other = 1; }
                else { p0 = 0; p1 = 2; p2 = 1; p3 = 3; other = 2; }
                const double dd[2] = {fcx(0, 0) - fcx(other, 0), fcx(0, 1) - fcx(other, 1)};
                const double mid = 2 * nrm(dd);
                const double h1 = 0.5 * (l[p0] + l[p1] + mid);
                v = std::sqrt(h1 * (h1 - l[p0]) * (h1 - l[p1]) * (h1 - mid));
                const double h2 = 0.5 * (l[p2] + l[p3] + mid);
                v += std::sqrt(h2 * (h2 - l[p2]) * (h2 - l[p3]) * (h2 - mid));
            }
            vol[c] = v;
        }
    }
    // orientation, eta, flags (MshBlock.cpp:281-305, Face.cpp:62-69)
#pragma omp parallel for schedule(static)
    for (int64_t f = 0; f < nfaces; f++) {
        double dot = 0.0;
        for (int d = 0; d < D; d++) dot += S[f * D + d] * (fc[f * D + d] - cc[(int64_t)c0[f] * D + d]);
        const int8_t o = dot < 0 ? -1 : 1;
        dac[f] = o;
        if (c1[f] >= 0) {
            double a[3], b[3];
            for (int d = 0; d < D; d++) {
                a[d] = cc[(int64_t)c0[f] * D + d] - fc[f * D + d];
                b[d] = cc[(int64_t)c1[f] * D + d] - fc[f * D + d];
            }
            const double d0 = nrm(a), d1 = nrm(b);
            eta[f] = d1 / (d0 + d1);
        } else {
            eta[f] = 1.0;
        }
        for (int d = 0; d < D; d++) {
            const bool out_ge0 = (double)o * S[f * D + d] >= 0;
            bool fl;
            if (flag_convention == 0) fl = out_ge0;
            else if (D == 2) fl = (o == -1) ? out_ge0 : !out_ge0;
            else fl = (o == -1) ? out_ge0 : false;
            flag[f * D + d] = fl ? 1 : 0;
        }
    }
    if (D == 3) {
        // extension: V = 1/3 sum_j Sout_j . (fc_j - cc)
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < ncells; c++) {
            double v = 0.0;
            for (int j = cf_ptr[c]; j < cf_ptr[c + 1]; j++) {
                const int64_t f = cf_idx[j];
                const double sg = ((c0[f] == c) ? 1.0 : -1.0) * (double)dac[f];
                double dot = 0.0;
                for (int d = 0; d < 3; d++) dot += S[f * 3 + d] * (fc[f * 3 + d] - cc[c * 3 + d]);
                v += sg * dot;
            }
            vol[c] = v / 3.0;
        }
    }
    return 0;
}

}  // extern "C"
