// mshread.cpp -- native reader / writer for the Fluent ASCII .msh subset that the
// reference's MshBlock accepts, producing the raw tables msthost_flatten consumes
// (at 50 M cells the reference's reader, not the solver, dominates the
// wall time -- it pushes every line through three std::stringstream copies and builds a
// pointer graph of std::vectors).  R = /root/reference/MST-CFD.
//
// What the reference does, and what is kept (the tables come out identical; pinned by
// tests/test_msh_reader_cpu.py against the digests of the reference build's own getters):
//   R/mesh/MshBlock.cpp:75-110   lines starting with '(' or ')' are header text, every other
//                                line is data; data lines are consumed strictly in file order
//   R/mesh/MshBlock.cpp:137-172  "(10 (0 first last ..." / "(12 (0 ..." / "(13 (0 ..." declare
//                                the node / cell / face counts (third token, hex)
//   R/mesh/MshBlock.cpp:176-190  a node zone reads DIM coordinates per line with stod()
//   R/mesh/MshBlock.cpp:191-262  a face zone "(13 (id first last type npf)(" reads npf node ids,
//                                c0 and -- for type 2 only -- c1 per line, all hex, 1-based;
//                                its name is the text of the last (0 "...") comment
//   R/work/FUNCTION.cpp:41-55    hexStringToInt: lower-case hex.  (It weights digits by string
//                                length, so on Linux the '\r' of the shipped CRLF files corrupts
//                                the last id of each line: '\r' is stripped here, as the
//                                Windows C runtime the reference was written on does.)
//   R/mesh/Node.cpp:13-15        Node::addNbFace: a node's faces in file order (msthost_node_faces),
//                                what the output path's node averaging walks (Work.cpp:287-295)
// Deliberate differences: the dimension comes from the file's "(2 d)" line (the reference's
// DIM is a macro), tokens may be separated by any run of blanks, upper-case hex digits are
// accepted (the reference silently drops them), and malformed input is an error with a
// message instead of undefined behaviour.
//
// Everything that touches every line (line index, node and face parsing, writing) is OpenMP code.
#include <omp.h>

#include "textout.h"

#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

thread_local std::string g_err;

struct Zone {
    int32_t id, start, end, type, npf;  // start/end: 0-based half-open face range
    std::string name;
};

struct Msh {
    int32_t dim = 0;
    int64_t nnodes = -1, ncells = -1, nfaces = -1, nint = 0;
    int32_t npf = 0;  // row stride of face_nodes = max nodes per face over the zones
    std::vector<double> nodes;
    std::vector<int32_t> face_nodes, c0, c1, ftype;
    std::vector<Zone> zones;
};

inline bool blank(char c) { return c == ' ' || c == '\t'; }

// hex token -> value; returns the position after the token or nullptr
inline const char* parse_hex(const char* p, const char* e, int64_t& v) {
    while (p < e && blank(*p)) p++;
    const char* s = p;
    int64_t x = 0;
    for (; p < e; p++) {
        const char c = *p;
        int d;
        if (c >= '0' && c <= '9') d = c - '0';
        else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
        else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
        else break;
        if (p - s >= 15) return nullptr;  // more than 15 hex digits cannot be an id or a count
        x = x * 16 + d;
    }
    if (p == s) return nullptr;
    v = x;
    return p;
}

inline const char* parse_double(const char* p, const char* e, double& v) {
    while (p < e && blank(*p)) p++;
    if (p < e && *p == '+') p++;
    auto r = std::from_chars(p, e, v);  // correctly rounded, like the stod() of MshBlock.cpp:186
    if (r.ec == std::errc()) return r.ptr;
    // anything from_chars refuses goes through strtod like the reference
    std::string tmp(p, e);
    char* end = nullptr;
    v = std::strtod(tmp.c_str(), &end);
    if (end == tmp.c_str()) return nullptr;
    return p + (end - tmp.c_str());
}

// Line index of the whole file.  start[i] = offset of line i, start[nlines] = size.
struct Lines {
    const char* base = nullptr;
    std::vector<int64_t> start;
    int64_t count() const { return (int64_t)start.size() - 1; }
    const char* b(int64_t i) const { return base + start[i]; }
    const char* e(int64_t i) const {  // end of the text: without '\n' and trailing '\r'
        const char* lo = b(i);
        const char* q = base + start[i + 1];
        if (q > lo && q[-1] == '\n') q--;
        while (q > lo && q[-1] == '\r') q--;
        return q;
    }
};

void index_lines(const char* buf, int64_t n, Lines& L) {
    L.base = buf;
    const int nt = std::max(1, omp_get_max_threads());
    std::vector<int64_t> cnt(nt + 1, 0);
    const int64_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        const int64_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
        int64_t c = 0;
        const char* p = buf + lo;
        while (p < buf + hi) {
            const char* q = (const char*)memchr(p, '\n', (size_t)(buf + hi - p));
            if (!q) break;
            c++;
            p = q + 1;
        }
        cnt[t + 1] = c;
    }
    for (int t = 0; t < nt; t++) cnt[t + 1] += cnt[t];
    const int64_t nl = cnt[nt];
    const bool tail = n > 0 && buf[n - 1] != '\n';  // last line without a newline
    L.start.assign((size_t)(nl + (tail ? 1 : 0) + 1), 0);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        const int64_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
        int64_t k = cnt[t] + 1;
        const char* p = buf + lo;
        while (p < buf + hi) {
            const char* q = (const char*)memchr(p, '\n', (size_t)(buf + hi - p));
            if (!q) break;
            L.start[k++] = (q + 1) - buf;
            p = q + 1;
        }
    }
    L.start[0] = 0;
    L.start.back() = n;
}

// text between the second '(' and the matching ')' of a header such as "(13 (d 1 22f9 2 2)("
bool second_brackets(const char* b, const char* e, const char*& ib, const char*& ie) {
    const char* p = b + 1;
    while (p < e && *p != '(') p++;
    if (p >= e) return false;
    ib = p + 1;
    const char* q = ib;
    while (q < e && *q != ')') q++;
    if (q >= e) return false;
    ie = q;
    return true;
}

bool fail(const std::string& s) {
    g_err = s;
    return false;
}

bool parse(const char* buf, int64_t n, Msh& m) {
    Lines L;
    index_lines(buf, n, L);
    const int64_t nl = L.count();
    // header lines in file order, data lines in file order (MshBlock.cpp:86-93)
    std::vector<int64_t> hdr, data;
    {
        std::vector<uint8_t> kind((size_t)nl);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nl; i++) {
            const char* b = L.b(i);
            const char* e = L.e(i);
            kind[i] = (b == e) ? 2 : ((*b == '(' || *b == ')') ? 1 : 0);
        }
        // compaction in parallel: per-thread counts, prefix, fill (27 M lines for a 12.6 M-tet file)
        const int nt = std::max(1, omp_get_max_threads());
        std::vector<int64_t> cd((size_t)nt + 1, 0), chh((size_t)nt + 1, 0);
        const int64_t chunk = (nl + nt - 1) / nt;
        int64_t bad_blank = -1;
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nt; t++) {
            const int64_t lo = std::min(nl, t * chunk), hi = std::min(nl, lo + chunk);
            int64_t a = 0, b = 0;
            for (int64_t i = lo; i < hi; i++) { a += kind[i] == 0; b += kind[i] == 1; }
            cd[(size_t)t + 1] = a;
            chh[(size_t)t + 1] = b;
        }
        for (int t = 0; t < nt; t++) { cd[(size_t)t + 1] += cd[(size_t)t]; chh[(size_t)t + 1] += chh[(size_t)t]; }
        data.resize((size_t)cd[(size_t)nt]);
        hdr.resize((size_t)chh[(size_t)nt]);
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nt; t++) {
            const int64_t lo = std::min(nl, t * chunk), hi = std::min(nl, lo + chunk);
            int64_t a = cd[(size_t)t], b = chh[(size_t)t];
            for (int64_t i = lo; i < hi; i++) {
                if (kind[i] == 0) data[(size_t)a++] = i;
                else if (kind[i] == 1) hdr[(size_t)b++] = i;
                else if (i > 0 && i + 1 < nl && kind[i - 1] == 0 && kind[i + 1] == 0) {
#pragma omp critical
                    if (bad_blank < 0 || i < bad_blank) bad_blank = i;
                }
            }
        }
        // a blank line between two data lines: the reference would take it for a data line (MshBlock.cpp:92)
        // and throw in stod()
        if (bad_blank >= 0) return fail("blank line inside a data block (line " + std::to_string(bad_blank + 1) + ")");
    }
    int64_t cur = 0;  // next unread data line
    int64_t node_done = 0, face_done = 0;
    std::string comment;
    for (int64_t h : hdr) {
        const char* b = L.b(h);
        const char* e = L.e(h);
        if (e - b < 4 || *b != '(') continue;  // every header read below looks at b[0..3]
        if (b[1] == '0' && (b[2] == ' ' || b[2] == '"')) {  // (0 "comment")
            const char* q0 = (const char*)memchr(b, '"', (size_t)(e - b));
            const char* q1 = q0 ? (const char*)memchr(q0 + 1, '"', (size_t)(e - q0 - 1)) : nullptr;
            comment = (q0 && q1) ? std::string(q0 + 1, q1) : std::string();
            continue;
        }
        if (b[1] == '2' && b[2] == ' ') {  // (2 dim)
            m.dim = b[3] - '0';
            if (m.dim != 2 && m.dim != 3) return fail("unsupported dimension line: " + std::string(b, e));
            continue;
        }
        if (b[1] != '1' || (b[2] != '0' && b[2] != '2' && b[2] != '3') || b[3] != ' ') continue;
        const char *ib, *ie;
        if (!second_brackets(b, e, ib, ie)) return fail("malformed section header: " + std::string(b, e));
        int64_t tok[5] = {0, 0, 0, 0, 0};
        int ntok = 0;
        for (const char* p = ib; ntok < 5;) {
            const char* q = parse_hex(p, ie, tok[ntok]);
            if (!q) break;
            ntok++;
            p = q;
        }
        if (ntok < 3) return fail("section header with fewer than 3 fields: " + std::string(b, e));
        const char sec = b[2];
        if (tok[0] == 0) {  // declaration: total count = third field (MshBlock.cpp:141-171)
            if (tok[2] >= (int64_t)1 << 31) return fail("count does not fit 32-bit ids");
            if (sec == '0') {
                if (!m.dim) return fail("node declaration before the (2 dim) line");
                m.nnodes = tok[2];
                m.nodes.assign((size_t)(m.nnodes * m.dim), 0.0);
            } else if (sec == '2') {
                m.ncells = tok[2];
            } else {
                m.nfaces = tok[2];
                m.c0.assign((size_t)m.nfaces, -1);
                m.c1.assign((size_t)m.nfaces, -1);
                m.ftype.assign((size_t)m.nfaces, 0);
            }
            continue;
        }
        if (sec == '0') {
            // node zone: the reference reads every remaining node here (MshBlock.cpp:180-189)
            if (m.nnodes < 0) return fail("node zone before the node declaration");
            const int64_t cnt = m.nnodes - node_done;
            if (cur + cnt > (int64_t)data.size()) return fail("file ends inside the node block");
            const int D = m.dim;
            int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
            for (int64_t k = 0; k < cnt; k++) {
                const int64_t li = data[cur + k];
                const char* p = L.b(li);
                const char* pe = L.e(li);
                for (int d = 0; d < D; d++) {
                    double v;
                    p = parse_double(p, pe, v);
                    if (!p) { bad = 1; break; }
                    m.nodes[(size_t)((node_done + k) * D + d)] = v;
                }
            }
            if (bad) return fail("unreadable coordinate in the node block");
            cur += cnt;
            node_done += cnt;
        } else if (sec == '3') {
            if (ntok < 5) return fail("face zone header needs (id first last type nodes-per-face): " + std::string(b, e));
            if (m.nfaces < 0 || m.ncells < 0 || m.nnodes < 0) return fail("face zone before the node / cell / face declarations");
            for (int i = 0; i < 5; i++)
                if (tok[i] < 0 || tok[i] >= (int64_t)1 << 31) return fail("face zone header field does not fit a 32-bit id: " + std::string(b, e));
            Zone z;
            z.id = (int32_t)tok[0];
            z.start = (int32_t)face_done;
            z.end = (int32_t)tok[2];
            z.type = (int32_t)tok[3];
            z.npf = (int32_t)tok[4];
            z.name = comment;
            if (tok[1] != face_done + 1)
                return fail("face zone " + std::to_string(z.id) + " does not start where the previous one ended "
                            "(the reference reads faces strictly in file order)");
            if (z.end > m.nfaces || z.end < z.start) return fail("face zone range outside the declared face count");
            if (z.npf < 2 || z.npf > 4) return fail("nodes per face must be 2, 3 or 4 (mixed zones are not read by the reference)");
            if (z.npf > m.npf) {  // widen the node table (quads after triangles)
                std::vector<int32_t> w((size_t)m.nfaces * z.npf, -1);
                for (int64_t f = 0; f < face_done; f++)
                    for (int k = 0; k < m.npf; k++) w[(size_t)f * z.npf + k] = m.face_nodes[(size_t)f * m.npf + k];
                m.face_nodes.swap(w);
                m.npf = z.npf;
            }
            const int64_t cnt = z.end - z.start;
            if (cur + cnt > (int64_t)data.size()) return fail("file ends inside face zone " + std::to_string(z.id));
            const int npf = z.npf, stride = m.npf;
            const bool interior = z.type == 2;  // only type-2 zones read a second cell (MshBlock.cpp:242-259)
            const int64_t nn = m.nnodes, nc = m.ncells;
            int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
            for (int64_t k = 0; k < cnt; k++) {
                const int64_t li = data[cur + k], f = z.start + k;
                const char* p = L.b(li);
                const char* pe = L.e(li);
                int64_t v;
                bool ok = true;
                for (int j = 0; j < npf && ok; j++) {
                    p = parse_hex(p, pe, v);
                    ok = p && v >= 1 && v <= nn;
                    if (ok) m.face_nodes[(size_t)f * stride + j] = (int32_t)(v - 1);
                }
                if (ok) {
                    p = parse_hex(p, pe, v);
                    ok = p && v >= 1 && v <= nc;
                    if (ok) m.c0[(size_t)f] = (int32_t)(v - 1);
                }
                if (ok && interior) {
                    p = parse_hex(p, pe, v);
                    ok = p && v >= 1 && v <= nc;
                    if (ok) m.c1[(size_t)f] = (int32_t)(v - 1);
                }
                if (!ok) bad = 1;
                m.ftype[(size_t)f] = z.type;
            }
            if (bad) return fail("unreadable or out-of-range id in face zone " + std::to_string(z.id));
            cur += cnt;
            face_done = z.end;
            if (interior) m.nint = z.end;  // MshBlock.cpp:213-216
            m.zones.push_back(std::move(z));
        }
    }
    if (m.dim == 0 || m.nnodes < 0 || m.ncells < 0 || m.nfaces < 0) return fail("missing (2 dim) or a node / cell / face declaration");
    if (node_done != m.nnodes) return fail("node block shorter than declared");
    if (face_done != m.nfaces) return fail("face zones cover " + std::to_string(face_done) + " of " + std::to_string(m.nfaces) + " declared faces");
    if (m.npf == 0) m.npf = m.dim;
    return true;
}

}  // namespace

extern "C" {

typedef struct msthost_msh msthost_msh;  // opaque: a parsed file

const char* msthost_last_error(void) { return g_err.c_str(); }

// Parse `path`.  0 = ok (the handle owns the tables until msthost_msh_free), < 0 = error.
int msthost_msh_read(const char* path, msthost_msh** out) {
    if (!path || !out) { g_err = "null argument"; return -1; }
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { g_err = std::string("cannot open ") + path; return -2; }  // MshBlock.cpp:79-82 prints and goes on
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); g_err = "fstat failed"; return -2; }
    if (st.st_size == 0) { close(fd); g_err = "empty file"; return -3; }
    // the file is parsed in place: mapped read-only, pages come in as the parser's threads touch them
    void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { g_err = std::string("cannot map ") + path; return -2; }
    madvise(map, (size_t)st.st_size, MADV_WILLNEED);
    Msh* m = new Msh();
    const bool ok = parse(static_cast<const char*>(map), (int64_t)st.st_size, *m);
    munmap(map, (size_t)st.st_size);
    if (!ok) { delete m; return -3; }
    *out = reinterpret_cast<msthost_msh*>(m);
    return 0;
}

// The same parser on a memory image of the file (hosts that already hold the text).
int msthost_msh_parse(const char* text, int64_t nbytes, msthost_msh** out) {
    if (!text || !out || nbytes < 0) { g_err = "null argument"; return -1; }
    *out = nullptr;
    Msh* m = new Msh();
    if (!parse(text, nbytes, *m)) { delete m; return -3; }
    *out = reinterpret_cast<msthost_msh*>(m);
    return 0;
}

void msthost_msh_free(msthost_msh* h) { delete reinterpret_cast<Msh*>(h); }

// sizes8 = dim, nnodes, ncells, nfaces, nint (MshBlock::getNumOfIntFaces), nzones, npf (row stride
// of face_nodes), 0
int msthost_msh_sizes(const msthost_msh* h, int64_t* sizes8) {
    if (!h || !sizes8) return -1;
    const Msh& m = *reinterpret_cast<const Msh*>(h);
    const int64_t s[8] = {m.dim, m.nnodes, m.ncells, m.nfaces, m.nint, (int64_t)m.zones.size(), m.npf, 0};
    memcpy(sizes8, s, sizeof s);
    return 0;
}

// Copies the tables out (any pointer may be null).  nodes [nnodes*dim]; face_nodes [nfaces*npf],
// 0-based, -1 padded; c0, c1 [nfaces] 0-based, c1 = -1 on boundary faces; ftype [nfaces];
// zones [nzones*5] = id, start (0-based), end (exclusive), type, nodes per face.
int msthost_msh_tables(const msthost_msh* h, double* nodes, int32_t* face_nodes, int32_t* c0, int32_t* c1,
                       int32_t* ftype, int32_t* zones) {
    if (!h) return -1;
    const Msh& m = *reinterpret_cast<const Msh*>(h);
    if (nodes) memcpy(nodes, m.nodes.data(), m.nodes.size() * sizeof(double));
    if (face_nodes) memcpy(face_nodes, m.face_nodes.data(), m.face_nodes.size() * sizeof(int32_t));
    if (c0) memcpy(c0, m.c0.data(), m.c0.size() * sizeof(int32_t));
    if (c1) memcpy(c1, m.c1.data(), m.c1.size() * sizeof(int32_t));
    if (ftype) memcpy(ftype, m.ftype.data(), m.ftype.size() * sizeof(int32_t));
    if (zones)
        for (size_t z = 0; z < m.zones.size(); z++) {
            const Zone& q = m.zones[z];
            const int32_t r[5] = {q.id, q.start, q.end, q.type, q.npf};
            memcpy(zones + 5 * z, r, sizeof r);
        }
    return 0;
}

const char* msthost_msh_zone_name(const msthost_msh* h, int32_t z) {  // FacesInf::getName
    const Msh& m = *reinterpret_cast<const Msh*>(h);
    return (z >= 0 && (size_t)z < m.zones.size()) ? m.zones[(size_t)z].name.c_str() : "";
}

// Node -> faces in the order Node::addNbFace builds it (R/mesh/Node.cpp:13-15 called from
// MshBlock.cpp:225-227: faces in file order, a face's nodes in line order).  nf_ptr [nnodes+1],
// nf_idx [sum of nodes per face].
int msthost_node_faces(int64_t nnodes, int64_t nfaces, int32_t npf, const int32_t* face_nodes, int32_t* nf_ptr,
                       int32_t* nf_idx) {
    if (!face_nodes || !nf_ptr || !nf_idx || npf < 1) return -1;
    std::vector<int32_t> cnt((size_t)nnodes + 1, 0);
    for (int64_t f = 0; f < nfaces; f++)
        for (int k = 0; k < npf; k++) {
            const int32_t v = face_nodes[f * npf + k];
            if (v >= 0) {
                if (v >= nnodes) return -1;
                cnt[(size_t)v + 1]++;
            }
        }
    nf_ptr[0] = 0;
    for (int64_t i = 0; i < nnodes; i++) nf_ptr[i + 1] = nf_ptr[i] + cnt[(size_t)i + 1];
    std::vector<int32_t> pos(nf_ptr, nf_ptr + nnodes);
    for (int64_t f = 0; f < nfaces; f++)
        for (int k = 0; k < npf; k++) {
            const int32_t v = face_nodes[f * npf + k];
            if (v >= 0) nf_idx[pos[(size_t)v]++] = (int32_t)f;
        }
    return 0;
}

// Cell -> nodes as MshBlock.cpp:335-368 derives it for the Tecplot connectivity list is NOT
// reproduced here (the element list is written by the host's own Work.cpp).

// Writes raw tables in the subset above (LF line ends, shortest round-trip decimals), so that a
// synthetic mesh can be handed to the reference's own reader.  zones [nzones*5] as in
// msthost_msh_tables; names may be null.
int msthost_msh_write(const char* path, int32_t dim, int64_t nnodes, int64_t ncells, int64_t nfaces, int32_t npf,
                      const double* nodes, const int32_t* face_nodes, const int32_t* c0, const int32_t* c1,
                      int32_t nzones, const int32_t* zones) {
    if (!path || !nodes || !face_nodes || !c0 || !c1 || !zones) { g_err = "null argument"; return -1; }
    FILE* fp = fopen(path, "wb");
    if (!fp) { g_err = std::string("cannot create ") + path; return -2; }
    fprintf(fp, "(0 \" written by msthost_msh_write\")\n(2 %d)\n(0 \"Node Section\")\n", dim);
    fprintf(fp, "(10 (0 1 %llx 0 %d))\n(10 (5 1 %llx 1 %d)\n(\n", (unsigned long long)nnodes, dim,
            (unsigned long long)nnodes, dim);
    using msthost::put_hex;
    bool wrote = true;
    auto emit = [&](int64_t count, int width, auto&& line) { wrote = msthost::emit_records(fp, count, width, line) && wrote; };
    emit(nnodes, 32 * dim, [&](int64_t i, char* p) {
        for (int d = 0; d < dim; d++) {
            if (d) *p++ = ' ';
            p = std::to_chars(p, p + 30, nodes[i * dim + d]).ptr;
        }
        *p++ = '\n';
        return p;
    });
    fprintf(fp, "))\n(12 (0 1 %llx 0 0))\n(12 (6 1 %llx 1 1))\n(13 (0 1 %llx 0 0))\n", (unsigned long long)ncells,
            (unsigned long long)ncells, (unsigned long long)nfaces);
    for (int32_t z = 0; z < nzones; z++) {
        const int32_t* q = zones + 5 * z;
        const int32_t zs = q[1], ze = q[2], type = q[3], zn = q[4];
        fprintf(fp, "(0 \"Faces of zone Z%d\")\n(13 (%x %x %x %x %x)(\n", z, z + 7, zs + 1, ze, type, zn);
        emit(ze - zs, 9 * (zn + 2) + 2, [&](int64_t k, char* p) {
            const int64_t f = zs + k;
            for (int j = 0; j < zn; j++) {
                p = put_hex(p, (uint32_t)(face_nodes[f * npf + j] + 1));
                *p++ = ' ';
            }
            p = put_hex(p, (uint32_t)(c0[f] + 1));
            *p++ = ' ';
            p = put_hex(p, (uint32_t)(c1[f] >= 0 ? c1[f] + 1 : 0));
            *p++ = '\n';
            return p;
        });
        fprintf(fp, ")\n)\n");
    }
    fprintf(fp, "(0 \"Zone Sections\")\n");
    const bool ok = (fclose(fp) == 0) && wrote;
    if (!ok) g_err = "write failed";
    return ok ? 0 : -2;
}

}  // extern "C"
