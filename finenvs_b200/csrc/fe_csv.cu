// fe_csv.cu — host side of the loader: the market-data CSV reader (no device code in this file).
//
// Replaces `pandas.read_csv(path, names=[Date, Time, Open, High, Low, Close, Volume])` of the reference's read_data()
// (finenvs/environments/time_series_env.py:80-88) for the one format the reference's data uses
//     MM/DD/YYYY,HH:MM[:SS],open,high,low,close,volume
// with a memory-mapped, multi-threaded single pass: the file is cut into one chunk per thread at line boundaries,
// every thread counts its records (fe_csv_open) and later parses them into the caller's arrays (fe_csv_read):
//     date_key   64-bit FNV-1a hash of the Date field's bytes — the reference treats Date as a STRING (`unique()` :98,
//                equality scans :141-152), so equal strings <=> equal days is all the loader needs;
//     sec_of_day seconds since midnight of the Time field (for between_time("9:30", "15:59"), :90-91);
//     ohlc       the four price fields as float64, converted exactly like pandas' default C-parser converter
//                (`precise_xstrtod`: at most 17 digits accumulated in a double, one exact power-of-ten scaling) so that the
//                staged prices are bit-identical to what the reference feeds its tensors (:169-177) — a correctly rounded
//                strtod differs from it on inputs with more than 17 significant digits.
// Anything outside that format (quotes, missing or non-numeric fields, header lines, other time formats, more or fewer
// than 7 fields) makes fe_csv_read return FE_ECSV and the Python loader falls back to pandas for that file.
#include "../../include/finenvs_b200.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cfloat>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

namespace {

struct CsvFile {
    const char *data = nullptr;
    size_t size = 0;
    std::vector<size_t> begin;  // chunk k = bytes [begin[k], begin[k+1]), both at line starts
    std::vector<int64_t> rows;  // records (non-empty lines) per chunk
};

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// pandas/_libs/src/parser/tokenizer.c precise_xstrtod() — the converter pandas.read_csv uses by default
// (float_precision None / "high") — with decimal '.', sci 'E', no thousands separator; the whole field must be consumed.
// At most 17 significant digits are accumulated in a double (further integer digits only raise the exponent, further
// decimals are dropped) and the value is scaled by ONE multiplication or division with an exact power of ten.
// Returns false where pandas would not produce a plain finite number from the field.
const double kPow10[] = {
    1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20,
    1e21, 1e22, 1e23, 1e24, 1e25, 1e26, 1e27, 1e28, 1e29, 1e30, 1e31, 1e32, 1e33, 1e34, 1e35, 1e36, 1e37, 1e38, 1e39, 1e40};
constexpr int kMaxPow10 = (int)(sizeof(kPow10) / sizeof(kPow10[0])) - 1;
bool pandas_strtod(const char *p, const char *end, double *out) {
    const int max_digits = 17;
    bool negative = false;
    if (p < end && (*p == '-' || *p == '+')) { negative = *p == '-'; ++p; }
    double number = 0.;
    int num_digits = 0, num_decimals = 0, exponent = 0;
    while (p < end && is_digit(*p)) {
        if (num_digits < max_digits) { number = number * 10. + (*p - '0'); ++num_digits; } else { ++exponent; }
        ++p;
    }
    if (p < end && *p == '.') {
        ++p;
        while (num_digits < max_digits && p < end && is_digit(*p)) { number = number * 10. + (*p - '0'); ++p; ++num_digits; ++num_decimals; }
        if (num_digits >= max_digits)
            while (p < end && is_digit(*p)) ++p;
        exponent -= num_decimals;
    }
    if (num_digits == 0) return false;
    if (negative) number = -number;
    if (p < end && (*p == 'E' || *p == 'e')) {
        ++p;
        bool eneg = false;
        if (p < end && (*p == '-' || *p == '+')) { eneg = *p == '-'; ++p; }
        if (p >= end || !is_digit(*p)) return false;
        int n = 0;
        while (p < end && is_digit(*p)) { if (n < 100000) n = n * 10 + (*p - '0'); ++p; }
        exponent += eneg ? -n : n;
    }
    if (p != end) return false;
    // prices: anything that needs a power of ten beyond 1e40 is left to pandas (its table goes to 1e308, with a two-step
    // path for subnormals)
    if (exponent > kMaxPow10 || exponent < -kMaxPow10) return false;
    if (exponent > 0) number *= kPow10[exponent];
    else number /= kPow10[-exponent];
    *out = number;
    return number - number == 0.0; // finite
}

// "H:MM", "HH:MM" or "HH:MM:SS" -> seconds since midnight
bool parse_time(const char *p, const char *end, int32_t *out) {
    int part[3] = {0, 0, 0}, np = 0, nd = 0;
    for (; p < end; ++p) {
        if (is_digit(*p)) {
            if (++nd > 2) return false;
            part[np] = part[np] * 10 + (*p - '0');
        } else if (*p == ':') {
            if (nd == 0 || ++np > 2) return false;
            nd = 0;
        } else {
            return false;
        }
    }
    if (np < 1 || nd == 0 || part[0] > 23 || part[1] > 59 || part[2] > 59) return false;
    *out = part[0] * 3600 + part[1] * 60 + part[2];
    return true;
}

// one record: [p, end) without the line terminator
bool parse_record(const char *p, const char *end, int64_t *date_key, int32_t *sec, double *ohlc) {
    const char *f[8];
    int nf = 0;
    f[nf++] = p;
    for (const char *q = p; q < end; ++q) {
        if (*q == '"') return false;
        if (*q == ',') {
            if (nf == 7) return false; // more than 7 fields
            f[nf++] = q + 1;
        }
    }
    if (nf != 7) return false;
    f[7] = end + 1;
    if (f[1] - 1 == f[0]) return false; // empty date
    uint64_t h = 1469598103934665603ull;
    for (const char *q = f[0]; q < f[1] - 1; ++q) { h ^= (unsigned char)*q; h *= 1099511628211ull; }
    *date_key = (int64_t)h;
    if (!parse_time(f[1], f[2] - 1, sec)) return false;
    for (int c = 0; c < 4; ++c)
        if (!pandas_strtod(f[2 + c], f[3 + c] - 1, ohlc + c)) return false;
    double vol; // pandas parses the column too: a file whose Volume is not numeric goes to the fallback
    return pandas_strtod(f[6], end, &vol);
}

// calls fn(line_begin, line_end) for every non-empty line of [from, to); line_end excludes "\n" and a preceding "\r"
template <typename F> void for_each_line(const char *data, size_t from, size_t to, F fn) {
    size_t pos = from;
    while (pos < to) {
        const char *nl = (const char *)memchr(data + pos, '\n', to - pos);
        size_t stop = nl ? (size_t)(nl - data) : to;
        size_t e = stop;
        if (e > pos && data[e - 1] == '\r') --e;
        if (e > pos) fn(data + pos, data + e);
        pos = stop + 1;
    }
}

} // namespace

extern "C" {

int fe_csv_open(const char *path, int32_t num_threads, void **handle, int64_t *num_rows) {
    if (!path || !handle || !num_rows) return FE_EINVAL;
    *handle = nullptr;
    *num_rows = 0;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return FE_EIO;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { close(fd); return FE_EIO; }
    CsvFile *cf = new (std::nothrow) CsvFile;
    if (!cf) { close(fd); return FE_EIO; }
    cf->size = (size_t)sb.st_size;
    if (cf->size > 0) {
        void *m = mmap(nullptr, cf->size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); delete cf; return FE_EIO; }
        cf->data = (const char *)m;
    }
    close(fd);
    int K = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
    if (K < 1) K = 1;
    if (K > 64) K = 64;
    if ((size_t)K > cf->size / (1 << 16) + 1) K = (int)(cf->size / (1 << 16) + 1); // >= 64 KB per thread
    cf->begin.assign(K + 1, cf->size);
    cf->begin[0] = 0;
    for (int k = 1; k < K; ++k) {
        size_t pos = cf->size / K * k;
        if (pos < cf->begin[k - 1]) pos = cf->begin[k - 1];
        const char *nl = pos < cf->size ? (const char *)memchr(cf->data + pos, '\n', cf->size - pos) : nullptr;
        cf->begin[k] = nl ? (size_t)(nl - cf->data) + 1 : cf->size;
    }
    cf->rows.assign(K, 0);
    std::vector<std::thread> th;
    for (int k = 0; k < K; ++k)
        th.emplace_back([cf, k] {
            int64_t n = 0;
            for_each_line(cf->data, cf->begin[k], cf->begin[k + 1], [&](const char *, const char *) { ++n; });
            cf->rows[k] = n;
        });
    for (auto &t : th) t.join();
    for (int k = 0; k < K; ++k) *num_rows += cf->rows[k];
    *handle = cf;
    return 0;
}

int fe_csv_read(void *handle, int64_t *date_key, int32_t *sec_of_day, double *ohlc) {
    CsvFile *cf = (CsvFile *)handle;
    if (!cf || !date_key || !sec_of_day || !ohlc) return FE_EINVAL;
    const int K = (int)cf->rows.size();
    std::vector<int64_t> first(K, 0);
    for (int k = 1; k < K; ++k) first[k] = first[k - 1] + cf->rows[k - 1];
    std::vector<int> bad(K, 0);
    std::vector<std::thread> th;
    for (int k = 0; k < K; ++k)
        th.emplace_back([=, &bad] {
            int64_t r = first[k];
            for_each_line(cf->data, cf->begin[k], cf->begin[k + 1], [&](const char *b, const char *e) {
                if (!bad[k] && !parse_record(b, e, date_key + r, sec_of_day + r, ohlc + 4 * r)) bad[k] = 1;
                ++r;
            });
        });
    for (auto &t : th) t.join();
    for (int k = 0; k < K; ++k)
        if (bad[k]) return FE_ECSV;
    return 0;
}

void fe_csv_close(void *handle) {
    CsvFile *cf = (CsvFile *)handle;
    if (!cf) return;
    if (cf->data) munmap((void *)cf->data, cf->size);
    delete cf;
}

} // extern "C"
