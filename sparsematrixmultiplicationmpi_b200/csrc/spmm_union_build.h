// spmm_union_build.h — host-side builder of the "union" tile layout (spmm_union.cu).
//
// Restructures the CSR rows the reference walks one by one
//   /root/reference "Source Code/SparseMatrixFatVectorMultiply.cpp":17-28
// into work items for the union kernel. Pure host code (no CUDA types) so the layout search can be
// exercised without a device.
//
//   block   R consecutive rows walked over the ascending union of their columns: one B row read from shared
//           memory feeds R accumulators from registers (an absent entry is a 0.0 value).
//   slot    what one 4-lane team walks: a block, or one of up to 8 segments of a long block (folded at the end).
//   item    8 slots = the work of one consumer warp; entries are stored step-major (the 8 teams of the warp
//           read one contiguous line per step).
//   chunk   a run of consecutive items walked by one CTA per k-tile; the builder replays it against a window of
//           NG groups of 4 B rows (one TMA gather4 each) kept in shared memory: a column the window holds is a
//           hit, the missing columns of an item are loaded in ascending groups of four into the least recently
//           used groups that no item in flight (the D most recent ones) reads.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace spmm
{

struct alignas(16) UItem // 32 bytes, read by the kernel
{
    unsigned long long blob_off; // byte offset of the item's blob in the blob array (16-byte aligned)
    unsigned bytes;              // blob bytes (multiple of 16)
    unsigned ring_off;           // byte offset of the blob inside the shared-memory blob ring
    int n_groups;                // gather4 loads of this item
    int load_begin;              // first entry of the item in gcols / gslot (= item index * maxg)
    int row0;                    // first row of the item
    int drain;                   // 1: every earlier item must have finished before this one is loaded (window full)
};

struct UnionParams
{
    int R = 2;           // rows per block (2 or 4)
    int KT = 32;         // k-tile width in doubles (16 or 32): a window row is KT*8 bytes
    int D = 8;           // items in flight (consumer warps + lookahead)
    int n_chunks = 1;    // chunks (CTAs per k-tile)
    int smem_bytes = 0;  // shared memory the ring and the window may use together
    int split_len = 48;  // blocks with more entries are cut into segments of about this length
    int slots = 4;       // teams per warp = slots per item (4: 8-lane teams, 8: 4-lane teams)
    int max_groups = 0;  // cap on the window groups (0 = what fits)
    int item_cost = 24;  // fixed cost of an item in steps (chunks are cut to equal sums of steps + item_cost)
};

struct UnionLayout
{
    UnionParams p;
    int n_rows = 0, n_items = 0, NG = 0, ring_bytes = 0;
    std::vector<UItem> items;
    std::vector<int> chunk_first; // n_chunks + 1
    std::vector<int> gcols;       // 4 per gather group; item i owns the groups [i*maxg, i*maxg + n_groups)
    std::vector<int> gslot;       // window group each gather fills
    int maxg = 0;                 // gather groups reserved per item (fixed stride: a producer can fetch an item's loads
                                  // without having read its descriptor)
    std::vector<unsigned char> blob;
    // statistics
    long long union_entries = 0, slot_steps = 0, staged_rows = 0, max_blob = 0, split_blocks = 0;
    int max_steps = 0, max_item_groups = 0, drains = 0;
    std::string error;
};

constexpr int U_MAX_STEPS = 2040;

// Blob of an item with SL slots: u16 len[SL] | u8 block[SL] (relative to the item's first block, 0xFF unused) |
// u8 segments[SL] (head: 1..SL, continuation: 0) | i32 row0 | i32 steps | u32 has_split | pad to a multiple of 16 |
// ids: ceil(steps/4) x [SL slots][4 x u16 window row] | values: steps x [SL slots][R doubles]
inline unsigned union_hdr_bytes(int SL) { return (unsigned)((SL * 4 + 12 + 15) & ~15); }
inline unsigned union_blob_bytes(int steps, int R, int SL)
{
    return union_hdr_bytes(SL) + (unsigned)(((steps + 3) / 4) * SL * 8 + steps * SL * 8 * R);
}

// 0 on success; on failure `out->error` says why (not sorted, rows too long, window or ring too small).
int build_union_layout(int n_rows, int n_cols, const int *rowptr, const int *colidx, const double *vals,
                       const UnionParams &p, UnionLayout *out);

} // namespace spmm
