// Text loader of the path's input side: replaces `np.loadtxt(file_path)` of PointCloud.read_from_file
// (/root/reference/pointCloudToolbox.py:51), which at 100 M points costs minutes once the curvature
// itself takes a tenth of a second.  Host code only: the file is
// memory-mapped, cut at line boundaries into one piece per thread, and every token is converted with
// std::from_chars -- correctly rounded like Python's float(), so the table equals np.loadtxt's bit for
// bit.  Format = what the reference's scans use: numbers separated by blanks or tabs (np.loadtxt's default delimiter), one row per
// line, '#' starts a comment, empty lines are skipped, every row has the same number of columns.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "pct_internal.h"

namespace pct {
namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped() {
        if (p && n) munmap(const_cast<char*>(p), n);
        if (fd >= 0) close(fd);
    }
    bool open_file(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) return true;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { n = 0; return false; }
        p = static_cast<const char*>(m);
        madvise(m, n, MADV_SEQUENTIAL);
        return true;
    }
};

inline bool is_sep(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// one line [b, e): number of values, optionally stored to out (cols of them); -1 on a malformed token
template <typename T>
long parse_line(const char* b, const char* e, T* out, long max_out) {
    long count = 0;
    while (b < e) {
        while (b < e && is_sep(*b)) ++b;
        if (b >= e || *b == '#') break;
        const char* t = b;
        if (*t == '+') ++t;  // from_chars takes no leading plus, float() does
        double v = 0.0;
        const std::from_chars_result r = std::from_chars(t, e, v);
        if (r.ec == std::errc::invalid_argument || r.ptr == t) return -1;
        if (r.ec == std::errc::result_out_of_range) {
            // float(): overflow gives inf, underflow 0 / denormal; redo with strtod for the exact value
            std::string tok(t, r.ptr);
            v = std::strtod(tok.c_str(), nullptr);
        }
        if (r.ptr < e && !is_sep(*r.ptr) && *r.ptr != '#') return -1;
        if (out && count < max_out) out[count] = (T)v;  // float: round to nearest even, like astype(float32)
        ++count;
        b = r.ptr;
    }
    return count;
}

struct Piece {
    size_t begin, end;  // byte range, begins at a line start
    long rows = 0;      // non-empty lines
    long first_row = 0;
    long bad_line = -1; // piece-local index of the first malformed / ragged line
};

}  // namespace
}  // namespace pct

using namespace pct;

namespace {

inline bool is_data_line(const char* b, const char* e) {
    while (b < e && is_sep(*b)) ++b;
    return b < e && *b != '#';
}

// first `want` tokens of a PLY body line as float(); the rest of the line is not looked at (ref utils.py:993-994)
template <typename T>
bool parse_prefix(const char* b, const char* e, T* out, int want) {
    for (int i = 0; i < want; ++i) {
        while (b < e && is_sep(*b)) ++b;
        if (b >= e) return false;
        const char* t = b;
        if (*t == '+') ++t;
        double v = 0.0;
        const std::from_chars_result r = std::from_chars(t, e, v);
        if (r.ec == std::errc::invalid_argument || r.ptr == t) return false;
        if (r.ec == std::errc::result_out_of_range) v = std::strtod(std::string(t, r.ptr).c_str(), nullptr);
        if (r.ptr < e && !is_sep(*r.ptr)) return false;
        out[i] = (T)v;
        b = r.ptr;
    }
    return true;
}

// cut the file at line boundaries, count the data lines of every piece in parallel
// (every_line: PLY body, where each line is a row; begin0: offset of the first byte of the body)
void count_rows(const Mapped& f, int threads, std::vector<Piece>& pieces, bool every_line = false, size_t begin0 = 0) {
    const size_t span = f.n - begin0;
    threads = (int)std::min<size_t>((size_t)std::max(threads, 1), std::max<size_t>(1, span >> 16));
    pieces.assign((size_t)threads, Piece());
    for (int t = 0; t < threads; ++t) {
        size_t b = begin0 + span * (size_t)t / (size_t)threads;
        if (t > 0 && b > begin0) {  // advance to the next line start
            const char* nl = static_cast<const char*>(memchr(f.p + b - 1, '\n', f.n - (b - 1)));
            b = nl ? (size_t)(nl - f.p) + 1 : f.n;
        }
        pieces[(size_t)t].begin = b;
        if (t > 0) pieces[(size_t)t - 1].end = b;
    }
    pieces.back().end = f.n;
    auto walk = [&](Piece& pc) {
        const char* p = f.p + pc.begin;
        const char* end = f.p + pc.end;
        long r = 0;
        while (p < end) {
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            const char* e = nl ? nl : end;
            r += (every_line || is_data_line(p, e)) ? 1 : 0;
            p = nl ? nl + 1 : end;
        }
        pc.rows = r;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back([&, t] { walk(pieces[(size_t)t]); });
    walk(pieces[0]);
    for (auto& th : pool) th.join();
    long total = 0;
    for (auto& pc : pieces) { pc.first_row = total; total += pc.rows; }
}

int default_threads(int threads) { return threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency()); }

}  // namespace

extern "C" {

int pct_text_shape(const char* path, int64_t* rows, int64_t* cols) {
    PCT_REQUIRE(path && rows && cols, "pct_text_shape: NULL argument");
    Mapped f;
    if (!f.open_file(path)) { set_error(std::string("pct_text_shape: cannot open ") + path); return PCT_ERR_INVALID_ARGUMENT; }
    std::vector<Piece> pieces;
    count_rows(f, default_threads(0), pieces);
    *rows = pieces.back().first_row + pieces.back().rows;
    *cols = 0;
    const char* p = f.p;
    const char* end = f.p + f.n;
    while (p < end) {  // columns of the first data line
        const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
        const char* e = nl ? nl : end;
        if (is_data_line(p, e)) {
            const long c = parse_line<double>(p, e, nullptr, 0);
            if (c < 0) { set_error("could not convert a token of the first row to float"); return PCT_ERR_INVALID_ARGUMENT; }
            *cols = c;
            break;
        }
        p = nl ? nl + 1 : end;
    }
    return PCT_OK;
}

}  // extern "C"

namespace {
template <typename T>
int text_load(const char* path, int64_t rows, int64_t cols, T* out, int threads) {
    PCT_REQUIRE(path && rows >= 0 && cols >= 0 && (out || rows * cols == 0), "pct_text_load: bad argument");
    Mapped f;
    if (!f.open_file(path)) { set_error(std::string("pct_text_load: cannot open ") + path); return PCT_ERR_INVALID_ARGUMENT; }
    std::vector<Piece> pieces;
    count_rows(f, default_threads(threads), pieces);
    const long total = pieces.back().first_row + pieces.back().rows;
    if (total != rows) { set_error("pct_text_load: the file holds " + std::to_string(total) + " rows"); return PCT_ERR_INVALID_ARGUMENT; }
    auto walk = [&](Piece& pc) {
        const char* p = f.p + pc.begin;
        const char* end = f.p + pc.end;
        long r = 0;
        while (p < end) {
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            const char* e = nl ? nl : end;
            if (is_data_line(p, e)) {
                const long c = parse_line(p, e, out + (pc.first_row + r) * cols, cols);
                if (c != cols && pc.bad_line < 0) pc.bad_line = r;
                ++r;
            }
            p = nl ? nl + 1 : end;
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < pieces.size(); ++t) pool.emplace_back([&, t] { walk(pieces[t]); });
    walk(pieces[0]);
    for (auto& th : pool) th.join();
    for (auto& pc : pieces)
        if (pc.bad_line >= 0) {
            set_error("the number of columns changed (or a token is not a number) at row " + std::to_string(pc.first_row + pc.bad_line + 1));
            return PCT_ERR_INVALID_ARGUMENT;
        }
    return PCT_OK;
}
}  // namespace

extern "C" {

int pct_text_load(const char* path, int64_t rows, int64_t cols, double* out, int threads) {
    return text_load(path, rows, cols, out, threads);
}

int pct_text_load_f32(const char* path, int64_t rows, int64_t cols, float* out, int threads) {
    return text_load(path, rows, cols, out, threads);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// PLY body reader: parse_ply of /root/reference/utils.py:979-1004 (skip to the line "end_header", then
// float() of the first three tokens of EVERY following line, rounded to float32).
// ---------------------------------------------------------------------------------------------------
namespace {

// offset of the first byte after the line whose stripped text is "end_header"; SIZE_MAX when absent
size_t ply_body_offset(const Mapped& f) {
    const char* p = f.p;
    const char* end = f.p + f.n;
    while (p < end) {
        const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
        const char* b = p;
        const char* e = nl ? nl : end;
        while (b < e && (is_sep(*b))) ++b;
        while (e > b && (is_sep(e[-1]))) --e;
        if (e - b == 10 && memcmp(b, "end_header", 10) == 0) return nl ? (size_t)(nl + 1 - f.p) : f.n;
        p = nl ? nl + 1 : end;
    }
    return SIZE_MAX;
}

}  // namespace

extern "C" {

int pct_ply_shape(const char* path, int64_t* rows, int64_t* body_offset) {
    PCT_REQUIRE(path && rows && body_offset, "pct_ply_shape: NULL argument");
    Mapped f;
    if (!f.open_file(path)) { set_error(std::string("pct_ply_shape: cannot open ") + path); return PCT_ERR_INVALID_ARGUMENT; }
    const size_t off = ply_body_offset(f);
    if (off == SIZE_MAX) { set_error("pct_ply_shape: no end_header line"); return PCT_ERR_INVALID_ARGUMENT; }
    std::vector<Piece> pieces;
    count_rows(f, default_threads(0), pieces, true, off);
    *rows = pieces.back().first_row + pieces.back().rows;
    *body_offset = (int64_t)off;
    return PCT_OK;
}

int pct_ply_load_f32(const char* path, int64_t body_offset, int64_t rows, float* out, int threads) {
    PCT_REQUIRE(path && body_offset >= 0 && rows >= 0 && (out || rows == 0), "pct_ply_load_f32: bad argument");
    Mapped f;
    if (!f.open_file(path)) { set_error(std::string("pct_ply_load_f32: cannot open ") + path); return PCT_ERR_INVALID_ARGUMENT; }
    PCT_REQUIRE((size_t)body_offset <= f.n, "pct_ply_load_f32: body offset beyond the end of the file");
    std::vector<Piece> pieces;
    count_rows(f, default_threads(threads), pieces, true, (size_t)body_offset);
    const long total = pieces.back().first_row + pieces.back().rows;
    if (total != rows) { set_error("pct_ply_load_f32: the body holds " + std::to_string(total) + " lines"); return PCT_ERR_INVALID_ARGUMENT; }
    auto walk = [&](Piece& pc) {
        const char* p = f.p + pc.begin;
        const char* end = f.p + pc.end;
        long r = 0;
        while (p < end) {
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            const char* e = nl ? nl : end;
            if (!parse_prefix(p, e, out + (pc.first_row + r) * 3, 3) && pc.bad_line < 0) pc.bad_line = r;
            ++r;
            p = nl ? nl + 1 : end;
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 1; t < pieces.size(); ++t) pool.emplace_back([&, t] { walk(pieces[t]); });
    walk(pieces[0]);
    for (auto& th : pool) th.join();
    for (auto& pc : pieces)
        if (pc.bad_line >= 0) {
            set_error("body line " + std::to_string(pc.first_row + pc.bad_line + 1) + " does not start with three numbers");
            return PCT_ERR_INVALID_ARGUMENT;
        }
    return PCT_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Writers on the output side of the path.  Rows are formatted by all host threads, a block of rows per
// thread and wave, and every block is written at its own file offset (pwrite) -- the per-line Python
// loop of utils.py:538-551 takes ~1 us per value.
//   * points PLY: save_points_to_ply (utils.py:963-976), np.savetxt with '%.6f %.6f %.6f'
//   * curvature PLY: utils.py:538-551, f'{x} {y} {z} {K} {H}' with numpy float32 scalars -- their
//     __format__ goes through Python's float, so every value is repr(float(v)): the shortest digits that
//     round-trip the DOUBLE image of the float32, fixed notation for 1e-4 <= |v| < 1e16
// ---------------------------------------------------------------------------------------------------
namespace {

// repr(float(v)) of CPython (Python/pystrtod.c format_float_short, mode 'r')
inline char* py_repr(char* o, double v) {
    if (v != v) { memcpy(o, "nan", 3); return o + 3; }
    if (std::signbit(v)) { *o++ = '-'; v = -v; }
    if (v > 1.7976931348623157e308) { memcpy(o, "inf", 3); return o + 3; }
    if (v == 0.0) { memcpy(o, "0.0", 3); return o + 3; }
    char t[40];
    const std::to_chars_result r = std::to_chars(t, t + sizeof t, v, std::chars_format::scientific);  // d[.ddd]e+XX, shortest
    char digits[24];
    int nd = 0;
    const char* q = t;
    for (; q < r.ptr && *q != 'e'; ++q)
        if (*q != '.') digits[nd++] = *q;
    int ex = 0;
    {
        const char* x = q + 1;
        const bool neg = *x == '-';
        ++x;
        for (; x < r.ptr; ++x) ex = ex * 10 + (*x - '0');
        if (neg) ex = -ex;
    }
    const int decpt = ex + 1;  // value = 0.d1d2... * 10^decpt
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            *o++ = '0'; *o++ = '.';
            for (int i = 0; i < -decpt; ++i) *o++ = '0';
            memcpy(o, digits, (size_t)nd); o += nd;
        } else if (decpt >= nd) {
            memcpy(o, digits, (size_t)nd); o += nd;
            for (int i = nd; i < decpt; ++i) *o++ = '0';
            *o++ = '.'; *o++ = '0';
        } else {
            memcpy(o, digits, (size_t)decpt); o += decpt;
            *o++ = '.';
            memcpy(o, digits + decpt, (size_t)(nd - decpt)); o += nd - decpt;
        }
    } else {
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, digits + 1, (size_t)(nd - 1)); o += nd - 1; }
        *o++ = 'e';
        int e10 = decpt - 1;
        *o++ = e10 < 0 ? '-' : '+';
        if (e10 < 0) e10 = -e10;
        if (e10 >= 100) { *o++ = (char)('0' + e10 / 100); e10 %= 100; *o++ = (char)('0' + e10 / 10); *o++ = (char)('0' + e10 % 10); }
        else { *o++ = (char)('0' + e10 / 10); *o++ = (char)('0' + e10 % 10); }
    }
    return o;
}

// '%.6f' of C (what np.savetxt applies to a float64)
inline char* fixed6(char* o, double v) {
    if (v != v) { memcpy(o, "nan", 3); return o + 3; }  // Python's '%.6f' never prints a sign for nan
    if (v > 1.7976931348623157e308) { memcpy(o, "inf", 3); return o + 3; }
    if (v < -1.7976931348623157e308) { memcpy(o, "-inf", 4); return o + 4; }
    return std::to_chars(o, o + 330, v, std::chars_format::fixed, 6).ptr;
}

// rows [0, n) formatted by `fmt(row, out) -> end`, at most `max_row` bytes each, appended to `fd` at `offset`.
// Blocks of rows are claimed in order from a counter; a block's file offset is the end of the block before it, so
// its owner waits until the previous owner has FORMATTED (not written) its block and published where it ends --
// a chained scan: formatting and writing of different blocks overlap and no thread waits at a barrier.
template <typename F>
int write_rows(int fd, size_t offset, int64_t n, size_t max_row, int threads, F fmt) {
    threads = default_threads(threads);
    const int64_t block = 1 << 14;
    const int64_t blocks = (n + block - 1) / block;
    if (blocks == 0) return PCT_OK;
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, blocks));
    std::unique_ptr<std::atomic<int64_t>[]> start(new std::atomic<int64_t>[(size_t)blocks + 1]);
    for (int64_t b = 0; b <= blocks; ++b) start[(size_t)b].store(-1, std::memory_order_relaxed);
    start[0].store((int64_t)offset, std::memory_order_release);
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        std::vector<char> buf((size_t)block * max_row);
        for (;;) {
            const int64_t b = next.fetch_add(1, std::memory_order_relaxed);
            if (b >= blocks) return;
            const int64_t r0 = b * block, r1 = std::min(n, r0 + block);
            char* o = buf.data();
            for (int64_t i = r0; i < r1; ++i) o = fmt(i, o);
            const size_t len = (size_t)(o - buf.data());
            int64_t at;
            while ((at = start[(size_t)b].load(std::memory_order_acquire)) < 0) std::this_thread::yield();
            start[(size_t)b + 1].store(at + (int64_t)len, std::memory_order_release);
            const char* p = buf.data();
            size_t left = len, pos = (size_t)at;
            while (left && !failed.load(std::memory_order_relaxed)) {
                const ssize_t w = pwrite(fd, p, left, (off_t)pos);
                if (w <= 0) { failed.store(1); break; }
                p += w; pos += (size_t)w; left -= (size_t)w;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    if (failed.load()) { set_error("write failed"); return PCT_ERR_INVALID_ARGUMENT; }
    return PCT_OK;
}

int open_with_header(const char* path, const std::string& header) {
    const int fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return -1;
    if (pwrite(fd, header.data(), header.size(), 0) != (ssize_t)header.size()) { close(fd); return -1; }
    return fd;
}

}  // namespace

extern "C" {

int pct_write_points_ply(const char* path, const void* points, int is_f64, int64_t n, int threads) {
    PCT_REQUIRE(path && n >= 0 && (points || n == 0), "pct_write_points_ply: bad argument");
    const std::string header = "ply\nformat ascii 1.0\nelement vertex " + std::to_string(n) +
                               "\nproperty float x\nproperty float y\nproperty float z\nend_header\n";
    const int fd = open_with_header(path, header);
    if (fd < 0) { set_error(std::string("pct_write_points_ply: cannot write ") + path); return PCT_ERR_INVALID_ARGUMENT; }
    const float* pf = static_cast<const float*>(points);
    const double* pd = static_cast<const double*>(points);
    const int rc = write_rows(fd, header.size(), n, 3 * 332, threads, [=](int64_t i, char* o) {
        for (int c = 0; c < 3; ++c) {
            o = fixed6(o, is_f64 ? pd[3 * i + c] : (double)pf[3 * i + c]);
            *o++ = c < 2 ? ' ' : '\n';
        }
        return o;
    });
    close(fd);
    return rc;
}

int pct_write_curvature_ply(const char* path, const float* points, const float* gaussian, const float* mean, int64_t n,
                            int threads) {
    PCT_REQUIRE(path && n >= 0 && ((points && gaussian && mean) || n == 0), "pct_write_curvature_ply: bad argument");
    const std::string header = "ply\nformat ascii 1.0\nelement vertex " + std::to_string(n) +
                               "\nproperty float x\nproperty float y\nproperty float z\n"
                               "property float gaussian_curvature\nproperty float mean_curvature\nend_header\n";
    const int fd = open_with_header(path, header);
    if (fd < 0) { set_error(std::string("pct_write_curvature_ply: cannot write ") + path); return PCT_ERR_INVALID_ARGUMENT; }
    const int rc = write_rows(fd, header.size(), n, 5 * 26, threads, [=](int64_t i, char* o) {
        o = py_repr(o, (double)points[3 * i]);     *o++ = ' ';
        o = py_repr(o, (double)points[3 * i + 1]); *o++ = ' ';
        o = py_repr(o, (double)points[3 * i + 2]); *o++ = ' ';
        o = py_repr(o, (double)gaussian[i]);       *o++ = ' ';
        o = py_repr(o, (double)mean[i]);           *o++ = '\n';
        return o;
    });
    close(fd);
    return rc;
}

}  // extern "C"
