// textout.h -- ordered multi-threaded text output: every thread formats a contiguous block of
// records into its own buffer, the buffers are written in record order.  Used by the .msh writer
// (mshread.cpp) and the Tecplot writer (pltwrite.cpp).
#pragma once
#include <omp.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace msthost {

// line(i, p) writes record i at p and returns the new end; no record may exceed `width` bytes.
template <class F>
bool emit_records(FILE* fp, int64_t count, int width, F&& line) {
    const int nt = std::max(1, omp_get_max_threads());
    const int64_t block = 1 << 15;
    std::vector<std::string> out((size_t)nt);
    bool ok = true;
    for (int64_t base = 0; base < count && ok; base += block * nt) {
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nt; t++) {
            const int64_t lo = std::min(count, base + t * block), hi = std::min(count, lo + block);
            std::string& s = out[(size_t)t];
            s.resize((size_t)((hi - lo) * width));
            char* p = s.data();
            for (int64_t i = lo; i < hi; i++) p = line(i, p);
            s.resize((size_t)(p - s.data()));
        }
        for (auto& s : out)
            if (!s.empty() && fwrite(s.data(), 1, s.size(), fp) != s.size()) ok = false;
    }
    return ok;
}

inline char* put_dec(char* p, uint32_t v) {
    char t[10];
    int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = t[--n];
    return p;
}

inline char* put_hex(char* p, uint32_t v) {
    char t[8];
    int n = 0;
    do { t[n++] = "0123456789abcdef"[v & 15]; v >>= 4; } while (v);
    while (n) *p++ = t[--n];
    return p;
}

}  // namespace msthost
