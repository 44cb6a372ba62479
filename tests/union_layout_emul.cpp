// CPU emulation of spmm_union_kernel driven by the layout csrc/spmm_union_build.cu produces: checks the layout
// (blobs, gather lists, window slots, in-flight rule) without a GPU. Built and run by tests/test_union_layout.py.
//
// For every chunk the items are "consumed" in order while the producer is allowed to run as far ahead as the kernel's
// barriers permit (item w may be loaded once every item <= w-D has finished; a drain item only after all earlier ones):
// if a load ever overwrote a window row an unfinished item still reads, the result would differ from the CSR multiply.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "spmm_union_build.h"

using namespace spmm;

static int check(int n, int per_row, int half_bw, int hub_every, int hub_len, bool dups, int R, int SL, int KT, int D, int chunks,
                 int smem)
{
    std::mt19937_64 rng(n * 31 + per_row);
    std::vector<int> rp(1, 0), ci;
    std::vector<double> va;
    for (int i = 0; i < n; ++i)
    {
        const int m = (hub_every && i % hub_every == 3) ? hub_len : (i % 11 == 7 ? 0 : per_row);
        std::vector<int> c;
        for (int j = 0; j < m; ++j)
        {
            long long x = i + (long long)(rng() % (2 * half_bw + 1)) - half_bw;
            x = std::max(0ll, std::min<long long>(n - 1, x));
            c.push_back((int)x);
        }
        std::sort(c.begin(), c.end());
        if (!dups)
            c.erase(std::unique(c.begin(), c.end()), c.end());
        for (int x : c)
        {
            ci.push_back(x);
            va.push_back(0.5 + (double)(rng() % 1000) / 1000.0);
        }
        rp.push_back((int)ci.size());
    }
    UnionParams p;
    p.R = R;
    p.slots = SL;
    p.KT = KT;
    p.D = D;
    p.n_chunks = chunks;
    p.smem_bytes = smem;
    UnionLayout L;
    if (build_union_layout(n, n, rp.data(), ci.data(), va.data(), p, &L))
    {
        printf("build failed: %s\n", L.error.c_str());
        return 2; // a layout that does not fit is a legal outcome, reported to the caller
    }
    std::vector<double> B(n), ref(n, 0.0), out(n, NAN);
    for (int i = 0; i < n; ++i)
        B[i] = 1 + (double)(rng() % 100);
    for (int i = 0; i < n; ++i)
        for (int j = rp[i]; j < rp[i + 1]; ++j)
            ref[i] += va[j] * B[ci[j]];
    const unsigned hdr = union_hdr_bytes(SL);
    std::vector<double> window((size_t)L.NG * 4);
    for (int c = 0; c < L.p.n_chunks; ++c)
    {
        const int first = L.chunk_first[c], last = L.chunk_first[c + 1];
        std::fill(window.begin(), window.end(), NAN);
        int issued = first; // items [first, issued) have been loaded
        for (int i = first; i < last; ++i)
        {
            // the producer runs ahead: everything the barriers allow while item i is the oldest unfinished one
            while (issued < last)
            {
                const UItem &it = L.items[issued];
                const int w = issued - first, oldest = i - first;
                const bool may = it.drain ? (oldest >= w) : (w - D + 1 <= oldest); // items < target finished
                if (!may)
                    break;
                if (it.load_begin != issued * L.maxg || it.n_groups > L.maxg)
                    return 3;
                for (int g = 0; g < it.n_groups; ++g)
                    for (int q = 0; q < 4; ++q)
                    {
                        const int slot = L.gslot[(size_t)it.load_begin + g], col = L.gcols[((size_t)it.load_begin + g) * 4 + q];
                        if (slot < 0 || slot >= L.NG || col < 0 || col >= n)
                            return 4;
                        window[(size_t)slot * 4 + q] = B[col];
                    }
                ++issued;
            }
            if (issued <= i)
                return 5; // item i itself was never loaded
            const UItem &it = L.items[i];
            const unsigned char *bl = L.blob.data() + it.blob_off;
            const unsigned short *len16 = (const unsigned short *)bl;
            const unsigned char *blk8 = bl + SL * 2, *seg8 = bl + SL * 3;
            const int *h32 = (const int *)(bl + SL * 4);
            const int row0 = h32[0], steps = h32[1];
            if (row0 != it.row0 || it.bytes != union_blob_bytes(steps, R, SL) || it.ring_off + it.bytes > (unsigned)L.ring_bytes)
                return 6;
            const unsigned short *ids = (const unsigned short *)(bl + hdr);
            const double *vv = (const double *)(bl + hdr + ((steps + 3) / 4) * SL * 8);
            std::vector<double> acc((size_t)SL * R, 0.0);
            for (int t = 0; t < SL; ++t)
                for (int q = 0; q < len16[t]; ++q)
                {
                    const int id = ids[(q >> 2) * SL * 4 + t * 4 + (q & 3)];
                    if (id >= L.NG * 4)
                        return 7;
                    for (int r = 0; r < R; ++r)
                        acc[(size_t)t * R + r] = std::fma(vv[((size_t)q * SL + t) * R + r], window[id], acc[(size_t)t * R + r]);
                }
            for (int t = 0; t < SL; ++t)
            {
                if (blk8[t] == 0xFF || seg8[t] == 0)
                    continue;
                for (int r = 0; r < R; ++r)
                {
                    double s = acc[(size_t)t * R + r];
                    for (int j = 1; j < seg8[t]; ++j)
                        s += acc[(size_t)(t + j) * R + r];
                    const int row = row0 + blk8[t] * R + r;
                    if (row < n)
                        out[row] = s;
                }
            }
        }
    }
    double worst = 0;
    for (int i = 0; i < n; ++i)
    {
        const double e = std::fabs(out[i] - ref[i]) / std::max(1e-300, std::fabs(ref[i]));
        if (!(e <= 1e-12) && !(ref[i] == 0.0 && out[i] == 0.0))
        {
            printf("row %d: %g vs %g\n", i, out[i], ref[i]);
            return 1;
        }
        worst = std::max(worst, std::isnan(e) ? 0.0 : e);
    }
    printf("n=%d R=%d SL=%d KT=%d D=%d chunks=%d: items=%d NG=%d drains=%d staged/N=%.2f worst=%.2e\n", n, R, SL, KT, D, L.p.n_chunks,
           L.n_items, L.NG, L.drains, (double)L.staged_rows / n, worst);
    return 0;
}

int main()
{
    int fails = 0, built = 0;
    const int smem = 232448 - 3072;
    for (int R : {2, 4})
        for (int SL : {4, 8})
            for (int KT : {16, 32})
                for (int D : {2, 6, 12})
                {
                    const int rcs[] = {check(3001, 9, 40, 0, 0, false, R, SL, KT, D, 7, smem),
                                       check(5000, 20, 300, 97, 700, false, R, SL, KT, D, 13, smem),
                                       check(2000, 14, 25, 50, 120, true, R, SL, KT, D, 1, smem),
                                       check(1500, 12, 600, 0, 0, false, R, SL, KT, D, 5, 24 * 1024)}; // small window: drains
                    for (int rc : rcs)
                    {
                        if (rc == 0)
                            ++built;
                        else if (rc != 2)
                        {
                            printf("FAIL rc=%d (R=%d SL=%d KT=%d D=%d)\n", rc, R, SL, KT, D);
                            ++fails;
                        }
                    }
                }
    printf("built %d layouts, %d failures\n", built, fails);
    return (fails || built < 40) ? 1 : 0;
}
