// Text loader of the path's input side: replaces `np.loadtxt(file_path)` of PointCloud.read_from_file
// (/root/reference/pointCloudToolbox.py:51), which at 100 M points costs minutes once the curvature
// itself takes a tenth of a second.  Host code only: the file is
// memory-mapped, cut at line boundaries into one piece per thread, and every token is converted with
// std::from_chars -- correctly rounded like Python's float(), so the table equals np.loadtxt's bit for
// bit.  Format = what the reference's scans use: numbers separated by blanks, tabs or commas, one row per
// line, '#' starts a comment, empty lines are skipped, every row has the same number of columns.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
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

// cut the file at line boundaries, count the data lines of every piece in parallel
void count_rows(const Mapped& f, int threads, std::vector<Piece>& pieces) {
    threads = (int)std::min<size_t>((size_t)std::max(threads, 1), std::max<size_t>(1, f.n >> 16));
    pieces.assign((size_t)threads, Piece());
    for (int t = 0; t < threads; ++t) {
        size_t b = f.n * (size_t)t / (size_t)threads;
        if (t > 0 && b > 0) {  // advance to the next line start
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
            r += is_data_line(p, e) ? 1 : 0;
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
