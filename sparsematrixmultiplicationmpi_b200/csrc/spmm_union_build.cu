// spmm_union_build.cu — host builder of the union tile layout (see spmm_union_build.h). Pure C++.
#include "spmm_union_build.h"

#include <algorithm>
#include <climits>
#include <cstring>

namespace spmm
{

namespace
{
struct Slot
{
    int blk;        // block index
    int begin, end; // entries of the block's union list this slot walks
    int nseg;       // head: segments of the block (1..8); continuation: 0
};
struct ItemPlan
{
    int slot_begin, slot_end; // into the slot vector
    int steps;
};
} // namespace

int build_union_layout(int n_rows, int n_cols, const int *rowptr, const int *colidx, const double *vals,
                       const UnionParams &p, UnionLayout *out)
{
    UnionLayout &L = *out;
    L = UnionLayout();
    L.p = p;
    L.n_rows = n_rows;
    const int R = p.R, SL = p.slots;
    if ((R != 2 && R != 4) || (SL != 4 && SL != 8) || (p.KT != 16 && p.KT != 32) || p.D < 1 || p.n_chunks < 1 || n_rows <= 0)
    {
        L.error = "union layout: bad parameters";
        return 1;
    }
    // ---- 1. union list of every block: ascending (column, occurrence) merge of its R rows
    const int n_blocks = (n_rows + R - 1) / R;
    std::vector<long long> bptr((size_t)n_blocks + 1, 0);
    std::vector<int> ucol;
    std::vector<double> uval; // R per entry
    ucol.reserve((size_t)rowptr[n_rows]);
    uval.reserve((size_t)rowptr[n_rows] * (size_t)R / 2 + 16);
    for (int b = 0; b < n_blocks; ++b)
    {
        long long ptr[4], end[4];
        for (int r = 0; r < R; ++r)
        {
            const int row = b * R + r;
            ptr[r] = row < n_rows ? rowptr[row] : 0;
            end[r] = row < n_rows ? rowptr[row + 1] : 0;
            for (long long j = ptr[r] + 1; j < end[r]; ++j)
                if (colidx[j] < colidx[j - 1])
                {
                    L.error = "union layout: a row is not sorted by column";
                    return 1;
                }
        }
        for (;;)
        {
            int c = INT_MAX;
            for (int r = 0; r < R; ++r)
                if (ptr[r] < end[r])
                    c = std::min(c, colidx[ptr[r]]);
            if (c == INT_MAX)
                break;
            if (c < 0 || c >= n_cols)
            {
                L.error = "union layout: column index outside the matrix";
                return 1;
            }
            ucol.push_back(c);
            for (int r = 0; r < R; ++r)
            {
                double v = 0.0;
                if (ptr[r] < end[r] && colidx[ptr[r]] == c)
                    v = vals[ptr[r]++];
                uval.push_back(v);
            }
        }
        bptr[b + 1] = (long long)ucol.size();
    }
    L.union_entries = (long long)ucol.size();

    // ---- 2. slots and items (8 slots each, blocks in order; a long block takes several slots of one item)
    std::vector<Slot> slots;
    std::vector<ItemPlan> plan;
    slots.reserve((size_t)n_blocks + 1024);
    {
        int used = 0;
        ItemPlan cur = {0, 0, 0};
        auto close = [&]() {
            cur.slot_end = (int)slots.size();
            plan.push_back(cur);
            cur = {(int)slots.size(), 0, 0};
            used = 0;
        };
        for (int b = 0; b < n_blocks; ++b)
        {
            const int len = (int)(bptr[b + 1] - bptr[b]);
            int nseg = 1;
            if (len > p.split_len + p.split_len / 4)
                nseg = std::min(SL, (len + p.split_len - 1) / p.split_len);
            const int seg = (len + nseg - 1) / std::max(1, nseg);
            if (seg > U_MAX_STEPS)
            {
                L.error = "union layout: a row block is too long (more than 8 x 2040 union entries)";
                return 1;
            }
            if (used + nseg > SL)
                close();
            if (nseg > 1)
                ++L.split_blocks;
            for (int s = 0; s < nseg; ++s)
            {
                Slot sl;
                sl.blk = b;
                sl.begin = std::min(len, s * seg);
                sl.end = std::min(len, (s + 1) * seg);
                sl.nseg = s == 0 ? nseg : 0;
                slots.push_back(sl);
                cur.steps = std::max(cur.steps, sl.end - sl.begin);
            }
            used += nseg;
            if (used == SL)
                close();
        }
        if (used)
            close();
    }
    const int n_items = (int)plan.size();
    L.n_items = n_items;

    // ---- 3. chunks of equal work (steps + a fixed cost per item)
    const int n_chunks = std::max(1, std::min(p.n_chunks, n_items));
    L.chunk_first.assign((size_t)n_chunks + 1, 0);
    {
        std::vector<long long> pre((size_t)n_items + 1, 0);
        for (int i = 0; i < n_items; ++i)
            pre[i + 1] = pre[i] + plan[i].steps + p.item_cost;
        for (int c = 1; c < n_chunks; ++c)
        {
            const long long want = pre[n_items] * c / n_chunks;
            int i = (int)(std::lower_bound(pre.begin(), pre.end(), want) - pre.begin());
            L.chunk_first[c] = std::max(L.chunk_first[c - 1], std::min(i, n_items));
        }
        L.chunk_first[n_chunks] = n_items;
    }
    L.p.n_chunks = n_chunks;

    // ---- 4. blob ring size: the largest D consecutive blobs of a chunk, plus room to wrap
    std::vector<unsigned> bytes((size_t)n_items);
    unsigned max_blob = 0;
    for (int i = 0; i < n_items; ++i)
    {
        bytes[i] = (union_blob_bytes(plan[i].steps, R, SL) + 127u) & ~127u; // ring pieces are 128-byte aligned
        max_blob = std::max(max_blob, bytes[i]);
        L.max_steps = std::max(L.max_steps, plan[i].steps);
        L.slot_steps += (long long)plan[i].steps * SL;
    }
    L.max_blob = max_blob;
    long long ring_need = 0;
    for (int c = 0; c < n_chunks; ++c)
    {
        long long sum = 0;
        for (int i = L.chunk_first[c]; i < L.chunk_first[c + 1]; ++i)
        {
            sum += bytes[i];
            if (i - p.D >= L.chunk_first[c])
                sum -= bytes[i - p.D];
            ring_need = std::max(ring_need, sum);
        }
    }
    ring_need += max_blob; // a blob that does not fit before the end of the ring starts over at 0
    ring_need = (ring_need + 1023) & ~1023ll;
    const long long group_bytes = 4ll * p.KT * 8;
    long long ng = ((long long)p.smem_bytes - ring_need) / group_bytes;
    if (p.max_groups > 0)
        ng = std::min<long long>(ng, p.max_groups);
    if (ng < 8 || ng * 4 > 65535)
    {
        L.error = "union layout: blob ring leaves no room for the window (" + std::to_string(ring_need) + " bytes of ring)";
        return 1;
    }
    const int NG = (int)ng;
    L.NG = NG;
    L.ring_bytes = (int)ring_need;

    // ---- 5. replay every chunk against the window; emit items, loads and blobs
    L.items.resize((size_t)n_items);
    unsigned long long blob_total = 0;
    for (int i = 0; i < n_items; ++i)
    {
        L.items[i].blob_off = blob_total;
        L.items[i].bytes = union_blob_bytes(plan[i].steps, R, SL);
        blob_total += (L.items[i].bytes + 15u) & ~15u;
    }
    L.blob.assign((size_t)blob_total + 16, 0);
    std::vector<int> col_slot((size_t)n_cols, -1), col_stamp((size_t)n_cols, -1);
    std::vector<int> g_last((size_t)NG), g_cols((size_t)NG * 4);
    std::vector<int> need, miss;
    for (int c = 0; c < n_chunks; ++c)
    {
        const int first = L.chunk_first[c], last = L.chunk_first[c + 1];
        std::fill(g_last.begin(), g_last.end(), INT_MIN / 2);
        std::fill(g_cols.begin(), g_cols.end(), -1);
        long long head = 0;
        for (int i = first; i < last; ++i)
        {
            const int li = i - first; // position in the chunk = the kernel's item counter
            UItem &it = L.items[i];
            const ItemPlan &pl = plan[i];
            it.drain = 0;
            it.row0 = slots[pl.slot_begin].blk * R;
            // blob ring: next free piece; items li-D+1 .. li-1 are still in flight
            if (head + bytes[i] > ring_need)
                head = 0;
            for (int j = std::max(first, i - p.D + 1); j < i; ++j)
            {
                const long long o = L.items[j].ring_off;
                if (head < o + bytes[j] && o < head + bytes[i])
                {
                    L.error = "union layout: blob ring too small";
                    return 1;
                }
            }
            it.ring_off = (unsigned)head;
            head += bytes[i];
            // distinct columns of the item
            need.clear();
            for (int s = pl.slot_begin; s < pl.slot_end; ++s)
                for (long long e = bptr[slots[s].blk] + slots[s].begin; e < bptr[slots[s].blk] + slots[s].end; ++e)
                    if (col_stamp[ucol[e]] != i)
                    {
                        col_stamp[ucol[e]] = i;
                        need.push_back(ucol[e]);
                    }
            miss.clear();
            for (int col : need)
            {
                if (col_slot[col] >= 0)
                    g_last[col_slot[col] >> 2] = li;
                else
                    miss.push_back(col);
            }
            std::sort(miss.begin(), miss.end());
            it.load_begin = (int)L.gslot.size();
            it.n_groups = (int)((miss.size() + 3) / 4);
            L.max_item_groups = std::max(L.max_item_groups, it.n_groups);
            for (int g = 0; g < it.n_groups; ++g)
            {
                // victim: least recently used group that no item in flight reads
                int v = -1;
                for (int s = 0; s < NG; ++s)
                    if (g_last[s] <= li - p.D && (v < 0 || g_last[s] < g_last[v]))
                        v = s;
                if (v < 0)
                {
                    // nothing free among the groups the items in flight leave alone: this item waits for all earlier ones
                    for (int s = 0; s < NG; ++s)
                        if (g_last[s] < li && (v < 0 || g_last[s] < g_last[v]))
                            v = s;
                    if (v < 0)
                    {
                        L.error = "union layout: one item needs more B rows than the window holds";
                        return 1;
                    }
                    if (!it.drain)
                        ++L.drains;
                    it.drain = 1;
                }
                for (int q = 0; q < 4; ++q)
                {
                    const int old = g_cols[(size_t)v * 4 + q];
                    if (old >= 0 && (col_slot[old] >> 2) == v)
                        col_slot[old] = -1;
                }
                for (int q = 0; q < 4; ++q)
                {
                    const int col = miss[std::min(miss.size() - 1, (size_t)g * 4 + q)];
                    g_cols[(size_t)v * 4 + q] = col;
                    if (col_slot[col] < 0)
                        col_slot[col] = v * 4 + q;
                    L.gcols.push_back(col);
                }
                g_last[v] = li;
                L.gslot.push_back(v);
            }
            L.staged_rows += 4ll * it.n_groups;
            // blob
            unsigned char *bl = L.blob.data() + it.blob_off;
            unsigned short *len16 = reinterpret_cast<unsigned short *>(bl);
            unsigned char *blk8 = bl + SL * 2, *seg8 = bl + SL * 3;
            int *hdr32 = reinterpret_cast<int *>(bl + SL * 4);
            const int steps = pl.steps, steps4 = (steps + 3) / 4;
            const unsigned hdr = union_hdr_bytes(SL);
            unsigned short *ids = reinterpret_cast<unsigned short *>(bl + hdr);
            double *vv = reinterpret_cast<double *>(bl + hdr + steps4 * SL * 8);
            const int blk0 = slots[pl.slot_begin].blk;
            int has_split = 0;
            for (int t = 0; t < SL; ++t)
            {
                const int s = pl.slot_begin + t;
                if (s >= pl.slot_end)
                {
                    len16[t] = 0;
                    blk8[t] = 0xFF;
                    seg8[t] = 1;
                    continue;
                }
                const Slot &sl = slots[s];
                len16[t] = (unsigned short)(sl.end - sl.begin);
                blk8[t] = (unsigned char)(sl.blk - blk0);
                seg8[t] = (unsigned char)sl.nseg;
                if (sl.nseg != 1)
                    has_split = 1;
                const long long e0 = bptr[sl.blk] + sl.begin;
                for (int q = 0; q < sl.end - sl.begin; ++q)
                {
                    ids[(q >> 2) * SL * 4 + t * 4 + (q & 3)] = (unsigned short)col_slot[ucol[e0 + q]];
                    for (int r = 0; r < R; ++r)
                        vv[((size_t)q * SL + t) * R + r] = uval[(size_t)(e0 + q) * R + r];
                }
            }
            hdr32[0] = it.row0;
            hdr32[1] = steps;
            hdr32[2] = has_split;
        }
        // forget the window of this chunk
        for (int s = 0; s < NG * 4; ++s)
            if (g_cols[s] >= 0)
                col_slot[g_cols[s]] = -1;
    }
    // fixed stride per item: the producers address an item's loads without a dependent read of its descriptor
    {
        const int maxg = std::max(4, L.max_item_groups);
        std::vector<int> sc((size_t)n_items * maxg * 4, 0), ss((size_t)n_items * maxg, 0);
        for (int i = 0; i < n_items; ++i)
        {
            UItem &it = L.items[i];
            for (int g = 0; g < it.n_groups; ++g)
            {
                for (int q = 0; q < 4; ++q)
                    sc[((size_t)i * maxg + g) * 4 + q] = L.gcols[((size_t)it.load_begin + g) * 4 + q];
                ss[(size_t)i * maxg + g] = L.gslot[(size_t)it.load_begin + g];
            }
            it.load_begin = i * maxg;
        }
        L.gcols.swap(sc);
        L.gslot.swap(ss);
        L.maxg = maxg;
    }
    return 0;
}

} // namespace spmm

#ifdef SPMM_UNION_PROBE
// C entry for layout experiments without a device (tools/union_layout_probe.py): statistics only.
extern "C" int spmm_union_layout_probe(int n_rows, int n_cols, const int *rowptr, const int *colidx, const double *vals,
                                       int R, int KT, int D, int n_chunks, int smem_bytes, int split_len, int slots,
                                       long long *stats /* 12 */, char *err, int err_cap)
{
    spmm::UnionParams p;
    p.R = R;
    p.KT = KT;
    p.D = D;
    p.n_chunks = n_chunks;
    p.smem_bytes = smem_bytes;
    p.split_len = split_len;
    p.slots = slots;
    spmm::UnionLayout L;
    const int rc = spmm::build_union_layout(n_rows, n_cols, rowptr, colidx, vals, p, &L);
    stats[0] = L.n_items;
    stats[1] = L.NG;
    stats[2] = L.ring_bytes;
    stats[3] = L.union_entries;
    stats[4] = L.slot_steps;
    stats[5] = L.staged_rows;
    stats[6] = L.max_blob;
    stats[7] = L.max_steps;
    stats[8] = L.max_item_groups;
    stats[9] = L.drains;
    stats[10] = (long long)L.blob.size();
    stats[11] = (long long)L.gslot.size();
    if (err && err_cap > 0)
    {
        strncpy(err, L.error.c_str(), (size_t)err_cap - 1);
        err[err_cap - 1] = 0;
    }
    return rc;
}
#endif
