// mm_read.cu — native MatrixMarket front end: the text half of readMatrixMarketFile
//   /root/reference "Source Code/utils.cpp":70-153
// (the CSR assembly half, :124-181, is csr_build.cu on the device). Pure host C++.
//
// The reference reads the body with `file >> row >> col [>> value]`, i.e. as a stream of whitespace-separated tokens
// in which line breaks mean nothing (:128-136). This reader keeps exactly that: the body is cut at whitespace into one
// piece per thread, a first pass counts the tokens of every piece, the prefix sum tells each piece which field of which
// record its first token is, and a second pass converts the tokens in parallel (std::from_chars: correctly rounded, the
// same double operator>> produces). Header rules (:84-105): every leading line that starts with '%' is a comment; one
// that contains "symmetric" makes the matrix symmetric (so "skew-symmetric" does too, no sign flip), one that contains
// "pattern" gives every record the value 1.0. The first other line holds "rows cols entries" (:108-109).
// Errors carry the reference's messages (:77, :114, :140).
#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "spmm_internal.h"

namespace
{
inline bool is_space(unsigned char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// operator>>(int) semantics for one token: optional sign, digits; anything else fails the stream
inline bool parse_int(const char *b, const char *e, int *out)
{
    if (b < e && *b == '+')
    {
        ++b;
        if (b < e && (*b == '+' || *b == '-'))
            return false; // "+-3": from_chars would take the minus; operator>> fails
    }
    const auto r = std::from_chars(b, e, *out);
    return r.ec == std::errc() && r.ptr == e;
}
inline bool parse_double(const char *b, const char *e, double *out)
{
    if (b < e && *b == '+')
    {
        ++b;
        if (b < e && (*b == '+' || *b == '-'))
            return false;
    }
    const auto r = std::from_chars(b, e, *out);
    if (r.ec == std::errc() && r.ptr == e)
        return true;
    // forms from_chars rejects but strtod / num_get accept are rare (hex floats): fall back
    std::string tmp(b, e);
    char *end = nullptr;
    *out = strtod(tmp.c_str(), &end);
    return end && *end == 0 && !tmp.empty();
}
} // namespace

using namespace spmm;

extern "C"
{

int spmm_mm_read(const char *path, int *n_rows, int *n_cols, long long *n_entries, int *symmetric, int **rows,
                 int **cols, double **vals)
{
    SPMM_REQUIRE(path && n_rows && n_cols && n_entries && symmetric && rows && cols && vals, "NULL argument");
    *rows = *cols = nullptr;
    *vals = nullptr;
    const std::string name(path);
    FILE *f = fopen(path, "rb");
    if (!f)
    {
        set_error("Unable to open file: " + name);
        return SPMM_ERR_INVALID;
    }
    std::vector<char> buf;
    {
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf.resize(sz > 0 ? (size_t)sz : 0);
        const size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
        buf.resize(got);
        fclose(f);
    }
    const char *p = buf.data(), *end = buf.data() + buf.size();
    // ---- header: comment lines, then the size line (getline loop of utils.cpp:84-105)
    bool sym = false, pattern = false, have_size = false;
    const char *size_b = nullptr, *size_e = nullptr;
    while (p < end)
    {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        if (p < le && *p == '%')
        {
            const std::string line(p, le);
            sym |= line.find("symmetric") != std::string::npos;
            pattern |= line.find("pattern") != std::string::npos;
            p = nl ? nl + 1 : end;
            continue;
        }
        size_b = p;
        size_e = le;
        have_size = true;
        p = nl ? nl + 1 : end;
        break;
    }
    if (!have_size)
    {
        set_error("Failed to read matrix dimensions from file: " + name);
        return SPMM_ERR_INVALID;
    }
    int dims[3] = {0, 0, 0}; // a failed extraction leaves 0 (C++11 num_get), like the reference's stringstream
    {
        const char *q = size_b;
        for (int i = 0; i < 3; ++i)
        {
            while (q < size_e && is_space((unsigned char)*q))
                ++q;
            const char *t = q;
            while (q < size_e && !is_space((unsigned char)*q))
                ++q;
            if (t == q || !parse_int(t, q, &dims[i]))
                break;
        }
    }
    const long long nnz = dims[2];
    *n_rows = dims[0];
    *n_cols = dims[1];
    *symmetric = sym ? 1 : 0;
    *n_entries = std::max(0ll, nnz);
    if (nnz <= 0)
        return SPMM_OK;
    SPMM_REQUIRE(dims[0] > 0 && dims[1] > 0, "MatrixMarket size line: rows and columns must be positive");
    const int per = pattern ? 2 : 3;
    const long long need = nnz * per;

    // ---- body: pieces cut at whitespace, token counts, then parallel conversion
    const size_t body = (size_t)(end - p);
    int nt = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, body >> 16));
    nt = std::max(1, std::min(nt, 64));
    std::vector<const char *> cut((size_t)nt + 1);
    cut[0] = p;
    cut[nt] = end;
    for (int i = 1; i < nt; ++i)
    {
        const char *c = p + body * (size_t)i / (size_t)nt;
        c = std::max(c, cut[i - 1]);
        while (c < end && !is_space((unsigned char)*c)) // never cut inside a token
            ++c;
        cut[i] = c;
    }
    std::vector<long long> count((size_t)nt + 1, 0);
    {
        std::vector<std::thread> th;
        for (int i = 0; i < nt; ++i)
            th.emplace_back([&, i] {
                long long n = 0;
                const char *q = cut[i], *e = cut[i + 1];
                while (q < e)
                {
                    while (q < e && is_space((unsigned char)*q))
                        ++q;
                    if (q < e)
                        ++n;
                    while (q < e && !is_space((unsigned char)*q))
                        ++q;
                }
                count[i + 1] = n;
            });
        for (auto &t : th)
            t.join();
    }
    for (int i = 0; i < nt; ++i)
        count[i + 1] += count[i];
    int *r = (int *)malloc(sizeof(int) * (size_t)nnz), *c = (int *)malloc(sizeof(int) * (size_t)nnz);
    double *v = (double *)malloc(sizeof(double) * (size_t)nnz);
    if (!r || !c || !v)
    {
        free(r);
        free(c);
        free(v);
        set_error("out of host memory for the MatrixMarket records");
        return SPMM_ERR_NOMEM;
    }
    std::vector<long long> bad((size_t)nt, -1); // first token (global index) a piece could not convert
    {
        std::vector<std::thread> th;
        for (int i = 0; i < nt; ++i)
            th.emplace_back([&, i] {
                long long tok = count[i];
                const char *q = cut[i], *e = cut[i + 1];
                while (q < e && tok < need)
                {
                    while (q < e && is_space((unsigned char)*q))
                        ++q;
                    if (q >= e)
                        break;
                    const char *t = q;
                    while (q < e && !is_space((unsigned char)*q))
                        ++q;
                    const long long rec = tok / per;
                    const int field = (int)(tok % per);
                    bool ok;
                    if (field == 0)
                    {
                        ok = parse_int(t, q, &r[rec]);
                        --r[rec]; // 1-based -> 0-based (:143-144)
                    }
                    else if (field == 1)
                    {
                        ok = parse_int(t, q, &c[rec]);
                        --c[rec];
                    }
                    else
                        ok = parse_double(t, q, &v[rec]);
                    if (!ok && bad[i] < 0)
                        bad[i] = tok;
                    ++tok;
                }
            });
        for (auto &t : th)
            t.join();
    }
    bool failed = count[nt] < need; // the stream runs dry before the last record (:138-141)
    for (int i = 0; i < nt && !failed; ++i)
        failed = bad[i] >= 0;
    if (!failed)
        for (long long i = 0; i < nnz; ++i)
            if (r[i] < 0 || r[i] >= dims[0] || c[i] < 0 || c[i] >= dims[1])
            {
                free(r);
                free(c);
                free(v);
                set_error("MatrixMarket record " + std::to_string(i + 1) + " lies outside the declared " +
                          std::to_string(dims[0]) + " x " + std::to_string(dims[1]) + " matrix: " + name);
                return SPMM_ERR_INVALID;
            }
    if (failed)
    {
        free(r);
        free(c);
        free(v);
        set_error("Failed to read data from file: " + name);
        return SPMM_ERR_INVALID;
    }
    if (pattern)
        std::fill(v, v + nnz, 1.0); // :130-133
    *rows = r;
    *cols = c;
    *vals = v;
    return SPMM_OK;
}

void spmm_mm_free(int *rows, int *cols, double *vals)
{
    free(rows);
    free(cols);
    free(vals);
}

int spmm_csr_from_matrix_market(int device, const char *path, spmm_csr_t *out)
{
    SPMM_REQUIRE(out != nullptr, "out is NULL");
    int nr = 0, nc = 0, sym = 0;
    long long ne = 0;
    int *r = nullptr, *c = nullptr;
    double *v = nullptr;
    int rc = spmm_mm_read(path, &nr, &nc, &ne, &sym, &r, &c, &v);
    if (rc)
        return rc;
    rc = spmm_csr_from_coo_host(device, nr, nc, ne, r, c, v, sym, out);
    spmm_mm_free(r, c, v);
    return rc;
}

} // extern "C"
