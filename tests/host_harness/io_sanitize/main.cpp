#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <random>
#include "pct_b200.h"
// TEST INFRASTRUCTURE: csrc/pct_io.cu built alone with AddressSanitizer + UBSan, fed random bit patterns and garbage files.
static std::string g_dir;
static const char* P(const char* name) { static std::string s[8]; static int i = 0; std::string& r = s[i++ & 7]; r = g_dir + name; return r.c_str(); }
int main(int argc, char** argv) {
    g_dir = argc > 1 ? argv[1] : "/tmp";
    std::mt19937_64 rng(5);
    // writers on random bit patterns
    const long n = 100001;
    std::vector<float> pts(3 * n), K(n), H(n);
    for (auto* v : {&pts, &K, &H}) for (auto& x : *v) { uint32_t b = (uint32_t)rng(); memcpy(&x, &b, 4); }
    if (pct_write_curvature_ply(P("/asan_c.ply"), pts.data(), K.data(), H.data(), n, 5)) return 1;
    std::vector<double> pd(3 * n);
    for (auto& x : pd) { uint64_t b = rng(); memcpy(&x, &b, 8); }
    pd[0] = 1.7976931348623157e308; pd[1] = -1.7976931348623157e308; pd[2] = 4.9e-324;
    if (pct_write_points_ply(P("/asan_p.ply"), pd.data(), 1, n, 3)) return 2;
    if (pct_write_points_ply(P("/asan_p32.ply"), pts.data(), 0, n, 0)) return 3;
    // readers
    int64_t rows = 0, off = 0, cols = 0;
    if (pct_ply_shape(P("/asan_c.ply"), &rows, &off)) return 4;
    std::vector<float> back(3 * rows);
    int rc = pct_ply_load_f32(P("/asan_c.ply"), off, rows, back.data(), 7);
    printf("ply rows %ld rc %d\n", (long)rows, rc);
    // text: fuzzed garbage files must fail cleanly or load
    for (int it = 0; it < 300; ++it) {
        std::string s;
        const int len = (int)(rng() % 4000);
        const char alphabet[] = "0123456789.eE+- \t\n\r#naif,x";
        for (int i = 0; i < len; ++i) s += alphabet[rng() % (sizeof(alphabet) - 1)];
        if (it % 3 == 0) s += "\n";
        FILE* f = fopen(P("/asan_t.txt"), "wb"); fwrite(s.data(), 1, s.size(), f); fclose(f);
        if (pct_text_shape(P("/asan_t.txt"), &rows, &cols) == 0 && rows * cols < 100000) {
            std::vector<double> t((size_t)(rows * cols) + 1);
            pct_text_load(P("/asan_t.txt"), rows, cols, t.data(), 1 + it % 4);
            std::vector<float> t32((size_t)(rows * cols) + 1);
            pct_text_load_f32(P("/asan_t.txt"), rows, cols, t32.data(), 1 + it % 4);
        }
        if (pct_ply_shape(P("/asan_t.txt"), &rows, &off) == 0) {
            std::vector<float> b3((size_t)rows * 3 + 1);
            pct_ply_load_f32(P("/asan_t.txt"), off, rows, b3.data(), 2);
        }
    }
    puts("asan run done");
    return 0;
}
