// k_fused.cu -- ParquetScan -> Filter -> HashAggregate as one pass over the encoded pages (rows P1 + F1 + A1-A4 of
// SURVEY.md section 8a; BASELINE config 2).  The reference pipelines the three operators batch by batch through
// ParquetExec::execute (datafusion-ext-plans/src/parquet_exec.rs:151-204), FilterExec (filter_exec.rs:200-224) and
// AggTable::process_input_batch (agg/agg_table.rs:99-135); every stage materialises Arrow arrays for the next one.  On
// the GPU that materialisation is the cost: 3.4 GB of decoded columns written and re-read per SF100 pass.  Here a warp
// takes one tile of FZ_TILE rows and keeps it on chip:
//
//   scout   (one warp per page)   definition levels -> batch-wide validity bitmap (built in shared memory, flushed with
//                                 coalesced stores); the page is cut at the global-row multiples of FZ_TILE into segments,
//                                 each with a checkpoint of the dictionary-index stream -- so tiles of columns whose pages
//                                 do not line up are still short lists of segments
//   dict    (tiny)                the filter's interval test is evaluated ONCE per dictionary entry -> 1 bit per entry
//   staged  (persistent warps)    tiles whose columns are single segments of regular index streams (or PLAIN pages): the packed
//                                 bytes of the tile go to shared memory with one bulk copy per column (cp.async.bulk + mbarrier);
//                                 the predicate column is unpacked for the lane's rows, key and arguments only for the rows that
//                                 pass; the selected rows update the accumulators with L2 reductions (RED).  See fz_staged_kernel.
//   tile    (one warp per tile)   the tiles the staged kernel leaves on a work list (a page boundary inside the tile, RLE runs,
//                                 PLAIN fallback pages): the general segment walk -- predicate columns to pass bits with ballots,
//                                 key and argument columns to shared-memory planes, then the same row loop.
//   merge   (per batch)           see below
//
// Aggregation in dictionary space: a dictionary-encoded group key is never looked up.  A row's slot is its dictionary
// INDEX (offset by the dictionary's base in a per-batch accumulator array), which removes the divergent 4-byte gather per
// row (the L1 wavefront limit, ~2 cycles per lane per SM: the 1.59 ms of round 1's ss_item_sk decode) from the row loop; after
// the batch one thread per dictionary entry folds its accumulators into the direct table at key - kmin.  PLAIN key pages
// address the direct table straight away.
//
// Roofline: the only DRAM traffic that scales with rows is the encoded bytes (measured: 1.67 GB per SF100 pass against 1.44 GB
// algorithmic), but the kernels are bound by unpacking variable-width bit streams, not by moving them: ~3,900 warp instructions
// per 1024-row tile, 48 % of the issue slots (DESIGN.md section 6c has the history of what moved that number and what did not --
// more loads in flight and fewer atomics did not).
#include "device_utils.cuh"
#include "kernels.h"
#include "parquet_dev.h"
#include "parquet_hybrid.cuh"

namespace auron {

#define LAUNCH_CHECK(ctx)            \
    do {                             \
        CUDA_OK(cudaGetLastError()); \
        launch_count(ctx);           \
    } while (0)

// ------------------------------------------------------------------------------------------------------------ scout
constexpr int FZ_SB_WORDS = 1024;   // per-warp shared bitmap: pages of up to 32,768 rows (larger pages go through global memory)

// number of set bits in [b, b + m) of a bitmap, m <= FZ_TILE; warp-cooperative (<= 33 words)
template <bool GLOBAL>
__device__ __forceinline__ int fz_popc_range(const uint32_t* bm, int64_t b, int m, unsigned lane) {
    int c = 0;
    if (m > 0) {
        const int64_t w0 = b >> 5, w1 = (b + m - 1) >> 5;
        for (int64_t w = w0 + lane; w <= w1; w += 32) {
            uint32_t x = GLOBAL ? __ldcg(bm + w) : bm[w];
            if (w == w0) x &= 0xffffffffu << (b & 31);
            if (w == w1) x &= 0xffffffffu >> (31 - ((b + m - 1) & 31));
            c += __popc(x);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(FULL_MASK, c, d);
    return c;
}

__device__ __forceinline__ void fz_scout_page(const FzScoutCol& C, int page_id, int (*s_exit_w)[33], int* s_entry_w, uint32_t* sb) {
    const unsigned lane = lane_id();
    const PqPage pg = C.pages[page_id];
    const int rows = pg.num_values;
    if (rows <= 0) return;
    const int64_t gr0 = pg.row_start;
    const bool has_def = C.max_def > 0 && pg.def_len > 0;
    const bool use_sb = rows <= FZ_SB_WORDS * 32;
    // ---- 1. definition levels -> validity bits of rows [gr0, gr0 + rows) in the batch bitmap
    if (C.max_def > 0) {
        if (has_def && use_sb) {
            const int nw = (rows + 31) >> 5;
            for (int w = lane; w < nw; w += 32) sb[w] = 0;
            __syncwarp();
            lvl_page_bits(pg.def_ptr, pg.def_len, rows, lane, sb, s_exit_w, s_entry_w, 0);
            __syncwarp();
            const int sh = (int)(gr0 & 31);
            const int64_t q0 = gr0 >> 5;
            const int nd = (sh + rows + 31) >> 5;
            for (int k = lane; k < nd; k += 32) {
                const uint32_t cur = k < nw ? sb[k] : 0u, prev = k > 0 ? sb[k - 1] : 0u;
                const uint32_t bits = sh ? ((cur << sh) | (prev >> (32 - sh))) : cur;
                const bool full = (k > 0 || sh == 0) && ((int64_t)(k + 1) * 32 <= (int64_t)sh + rows);
                if (full) C.valid[q0 + k] = bits;
                else if (bits) atomicOr(&C.valid[q0 + k], bits);   // first / last word may be shared with the neighbouring page
            }
        } else if (has_def) {
            lvl_page_bits(pg.def_ptr, pg.def_len, rows, lane, C.valid, s_exit_w, s_entry_w, gr0);
            __threadfence();   // bits ORed by other lanes are counted below (read back through L2)
            __syncwarp();
        } else if (!pg.all_null) {   // no level section: every row holds a value
            const int64_t w0 = gr0 >> 5, w1 = (gr0 + rows - 1) >> 5;
            for (int64_t w = w0 + lane; w <= w1; w += 32) {
                uint32_t bits = 0xffffffffu;
                if (w == w0) bits &= 0xffffffffu << (gr0 & 31);
                if (w == w1) bits &= 0xffffffffu >> (31 - ((gr0 + rows - 1) & 31));
                if (w != w0 && w != w1) C.valid[w] = bits;
                else atomicOr(&C.valid[w], bits);
            }
        }
    }
    // ---- 2. segments: cut at the global-row multiples of FZ_TILE, checkpoint the index stream at every cut
    const bool dict = pg.encoding == 2 || pg.encoding == 8;
    const uint8_t* vals = pg.val_ptr;
    const uint8_t* idx_base = vals + 1;
    Hybrid idx;
    if (dict) idx.init(idx_base, vals + pg.val_len, pg.val_len > 0 ? vals[0] : 0);
    else idx.init(vals, vals, 0);
    const uint8_t* pf_idx = idx_base;
    const PqDict dd = dict ? C.dicts[pg.dict_id] : PqDict{nullptr, 0, 0};
    int seg = C.seg_base[page_id];
    int64_t v0 = 0;
    // ---- 2a. all segments at once.  Writers emit the indices of a high-cardinality column as maximal bit-packed runs (63 groups =
    // 504 values behind a one-byte header 0x7F; parquet-cpp and parquet-mr both close a literal run there), so the stream position
    // of any value is arithmetic -- once every run header has been checked to sit where that arithmetic puts it (induction over
    // the runs: a verified header fixes the start of the next run).  Lane s then builds segment s by itself (32 segments per
    // round): its non-null count from the bitmap, its first value from a warp scan, its checkpoint from the value number.
    // Anything else (RLE runs, short literal runs) takes the serial walk below.
    const int first_len = min(rows, FZ_TILE - (int)(gr0 & (FZ_TILE - 1)));
    const int nseg = 1 + (rows - first_len + FZ_TILE - 1) / FZ_TILE;
    {
        const int bw = dict ? (pg.val_len > 0 ? (int)vals[0] : 0) : -1;
        const int run_bytes = 1 + 63 * bw;
        // non-null values of the page (the stream's length in values) -> where its runs must sit
        int total_valid = rows;
        if (C.max_def > 0) {
            if (has_def) total_valid = 0;
            else if (pg.all_null) total_valid = 0;
        }
        if (C.max_def > 0 && has_def) {
            int c = 0;
            if (use_sb) {
                const int nw = (rows + 31) >> 5;   // (bits behind `rows` are zero)
                for (int w = lane; w < nw; w += 32) c += __popc(sb[w]);
            } else {
                const int64_t w0 = gr0 >> 5, w1 = (gr0 + rows - 1) >> 5;
                for (int64_t w = w0 + lane; w <= w1; w += 32) {
                    uint32_t x = __ldcg(C.valid + w);
                    if (w == w0) x &= 0xffffffffu << (gr0 & 31);
                    if (w == w1) x &= 0xffffffffu >> (31 - ((gr0 + rows - 1) & 31));
                    c += __popc(x);
                }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(FULL_MASK, c, d);
            total_valid = c;
        }
        const int nfull = total_valid / 504, rem = total_valid % 504, g_last = (rem + 7) / 8;
        bool regular = true;
        if (dict) {
            const int64_t ilen = (int64_t)pg.val_len - 1;
            regular = bw > 0 && bw <= 32 && (int64_t)nfull * run_bytes + (rem ? 1 + (int64_t)g_last * bw : 0) <= ilen;
            if (regular) {
                for (int j = lane; j < nfull; j += 32) regular = regular && idx_base[(int64_t)j * run_bytes] == 0x7F;
                if (rem && lane == 0) regular = regular && idx_base[(int64_t)nfull * run_bytes] == (uint8_t)((g_last << 1) | 1);
            }
            regular = __all_sync(FULL_MASK, regular);
        }
        if (regular) {
            int carry = 0;   // non-null values of the segments before this round
            for (int s0 = 0; s0 < nseg; s0 += 32) {
                const int sg = s0 + (int)lane;
                const int r_s = sg == 0 ? 0 : first_len + (sg - 1) * FZ_TILE;
                const int m_s = sg < nseg ? min(rows - r_s, sg == 0 ? first_len : FZ_TILE) : 0;
                int nv_s = m_s;
                if (C.max_def > 0) {
                    if (has_def) {
                        nv_s = 0;
                        if (m_s > 0) {
                            const int64_t b = use_sb ? (int64_t)r_s : gr0 + r_s;
                            const int64_t w0 = b >> 5, w1 = (b + m_s - 1) >> 5;
                            for (int64_t w = w0; w <= w1; w++) {
                                uint32_t x = use_sb ? sb[w] : __ldcg(C.valid + w);
                                if (w == w0) x &= 0xffffffffu << (b & 31);
                                if (w == w1) x &= 0xffffffffu >> (31 - ((b + m_s - 1) & 31));
                                nv_s += __popc(x);
                            }
                        }
                    } else if (pg.all_null) nv_s = 0;
                }
                int inc = nv_s;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(FULL_MASK, inc, d);
                    if ((int)lane >= d) inc += t;
                }
                const int v_s = carry + inc - nv_s;
                carry += __shfl_sync(FULL_MASK, inc, 31);
                if (sg < nseg) {
                    FzSeg S;
                    S.page = page_id;
                    S.row0 = r_s;
                    S.n = m_s;
                    S.nvalid = nv_s;
                    S.v0 = v_s;
                    S.idx = HybridCk{0, 0, 0, 0, 0, 1};
                    if (dict) {
                        const int j = v_s / 504, r = v_s % 504;
                        const int groups = j < nfull ? 63 : g_last;
                        const int hdr = j * run_bytes;
                        if (r == 0) S.idx = HybridCk{hdr, 0, hdr, 0, 0, 0};                                  // at a run boundary: the header is read on demand
                        else S.idx = HybridCk{hdr + 1 + groups * bw, groups * 8 - r, hdr + 1, r, 0, 0};      // inside run j
                    }
                    S.vals = vals;
                    S.val_len = pg.val_len;
                    S.bw = bw;
                    S.ddata = dict ? dd.data : nullptr;
                    S.ndict = dict ? dd.num_values : 0;
                    S.dict_id = pg.dict_id;
                    S.pad[0] = 1;
                    S.pad[1] = 0;
                    C.segs[seg + sg] = S;
                    if (((gr0 + r_s) & (FZ_TILE - 1)) == 0) C.first_seg[(gr0 + r_s) / FZ_TILE] = seg + sg;
                }
            }
            return;
        }
    }
    for (int r = 0; r < rows; seg++) {
        const int m = min(rows - r, FZ_TILE - (int)((gr0 + r) & (FZ_TILE - 1)));
        int nvalid = m;
        if (C.max_def > 0) {
            if (has_def) nvalid = use_sb ? fz_popc_range<false>(sb, r, m, lane) : fz_popc_range<true>(C.valid, gr0 + r, m, lane);
            else if (pg.all_null) nvalid = 0;
        }
        if (lane == 0) {
            FzSeg S;
            S.page = page_id;
            S.row0 = r;
            S.n = m;
            S.nvalid = nvalid;
            S.v0 = v0;
            S.idx = dict ? hybrid_save(idx, idx_base) : HybridCk{0, 0, 0, 0, 0, 1};
            S.vals = vals;
            S.val_len = pg.val_len;
            S.bw = dict ? (pg.val_len > 0 ? (int32_t)vals[0] : 0) : -1;
            S.ddata = dict ? dd.data : nullptr;
            S.ndict = dict ? dd.num_values : 0;
            S.dict_id = pg.dict_id;
            S.pad[0] = S.pad[1] = 0;
            C.segs[seg] = S;
            if (((gr0 + r) & (FZ_TILE - 1)) == 0) C.first_seg[(gr0 + r) / FZ_TILE] = seg;
        }
        if (dict) {   // the header walk is a pointer chase: pull the next 8 KB of the stream into L1 ahead of it
            const uint8_t* send = vals + pg.val_len;
            while (pf_idx < idx.p + 8192 && pf_idx < send) {
                const uint8_t* q = pf_idx + 128 * lane;
                if (q < send) asm volatile("prefetch.global.L1 [%0];" ::"l"(q));
                pf_idx += 4096;
            }
            hybrid_skip(idx, nvalid);
        }
        v0 += nvalid;
        r += m;
    }
}
__global__ void __launch_bounds__(PQ_WARPS * 32) fz_scout_kernel(const FzScoutCol* __restrict__ cols, const int32_t* __restrict__ page_base, int ncols) {
    __shared__ int s_exit[PQ_WARPS][32][33];
    __shared__ int s_entry[PQ_WARPS][33];
    __shared__ uint32_t s_bits[PQ_WARPS][FZ_SB_WORDS];
    const int wid = threadIdx.x >> 5;
    const int g = blockIdx.x * PQ_WARPS + wid;
    if (g >= page_base[ncols]) return;
    int c = 0;
    while (c + 1 < ncols && g >= page_base[c + 1]) c++;
    fz_scout_page(cols[c], g - page_base[c], s_exit[wid], s_entry[wid], s_bits[wid]);
}
void fz_scout(Ctx& ctx, const std::vector<FzScoutCol>& cols) {
    std::vector<int32_t> base{0};
    for (auto& c : cols) base.push_back(base.back() + c.n_pages);
    if (base.back() == 0) return;
    Buf dc = to_device(ctx, cols.data(), cols.size() * sizeof(FzScoutCol));
    Buf db = to_device(ctx, base.data(), base.size() * 4);
    ProfScope ps(ctx, "fz_scout");
    fz_scout_kernel<<<(base.back() + PQ_WARPS - 1) / PQ_WARPS, PQ_WARPS * 32, 0, ctx.stream>>>(P<FzScoutCol>(dc), P<int32_t>(db), (int)cols.size());
    LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------------------ dictionary predicate
// word w of pass_bits belongs to the dictionary d with pass_off[d] <= w < pass_off[d + 1]; bit i of it = entry 32 (w - pass_off[d]) + i
__global__ void __launch_bounds__(256) fz_dict_pass_kernel(const PqDict* __restrict__ dicts, const int32_t* __restrict__ pass_off, int n_dicts, int total_words,
                                                           int64_t lo, int64_t hi, uint32_t* __restrict__ pass_bits) {
    const int w = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5);
    if (w >= total_words) return;
    int a = 0, b = n_dicts - 1;
    while (a < b) {   // last dictionary whose first word is <= w
        const int m = (a + b + 1) >> 1;
        if (pass_off[m] <= w) a = m;
        else b = m - 1;
    }
    const PqDict d = dicts[a];
    const int i = (w - pass_off[a]) * 32 + (int)lane_id();
    bool pass = false;
    if (i < d.num_values) {
        const int64_t v = (int64_t)(int32_t)ld_u32_unaligned(d.data + (int64_t)i * 4);
        pass = v >= lo && v <= hi;
    }
    const uint32_t bits = __ballot_sync(FULL_MASK, pass);
    if (lane_id() == 0) pass_bits[w] = bits;
}
void fz_dict_pass(Ctx& ctx, const PqDict* dicts, const int32_t* pass_off, int n_dicts, int total_words, int64_t lo, int64_t hi, uint32_t* pass_bits) {
    if (total_words <= 0 || n_dicts <= 0) return;
    fz_dict_pass_kernel<<<(unsigned)(((int64_t)total_words * 32 + 255) / 256), 256, 0, ctx.stream>>>(dicts, pass_off, n_dicts, total_words, lo, hi, pass_bits);
    LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------------------ fused kernel
// One warp per tile of FZ_TILE rows; lane L owns rows [32 L, 32 L + 32) of the tile: its validity word, the number of
// non-null values before it (warp scan) and its selection word live in registers, so the row loop needs no ballots and no
// broadcast reads.  Columns are unpacked in VALUE order (rank = position among the non-null values of the tile), lanes
// striding over each run for coalesced loads:
//   predicate columns -> one PASS BIT per value in a 32-word rank-space bitmap (dictionary pages: a lookup in the per-entry
//                        bits of fz_dict_pass; PLAIN pages: the interval test), then every lane deposits its bits into the
//                        positions of its valid rows;
//   key column        -> accumulator slot per value (dictionary pages: slot base of the dictionary + index -- no lookup);
//   argument columns  -> value per value (dictionary lookup at unpack time).
// A lane then walks the set bits of its selection word.
//
// Latency: a tile's page bytes are a few KB per column, first touched by this warp.  Unpacking them step by step paid one DRAM
// round trip per step (35 % of all stall samples sat on the funnel shift behind those loads); instead the warp asks for ALL lines
// of ALL columns of the tile at once (prefetch.global.L2) before it starts on the first column.
constexpr int FZ_WARPS = 4;
constexpr int FZ_VSTRIDE = FZ_TILE + 32;    // value planes are padded by one word per 32 (rank r lives at r + r / 32): lanes that
                                            // read ranks 32 apart (columns without NULLs) hit different banks
__device__ __forceinline__ int fz_pi(int r) { return r + (r >> 5); }

struct FzX {                // what the transform of one segment needs
    bool dict, ddata_aligned;
    uint32_t ndict;
    const uint8_t* ddata;
    const uint32_t* pass;
    uint32_t slot_base;
    int64_t lo, hi;
    long long kmin;
    int64_t range;
    int32_t* oor;
};
template <int ROLE>
__device__ __forceinline__ uint32_t fz_xform(const FzX& x, uint32_t raw) {
    if (x.dict) {
        const uint32_t i = raw < x.ndict ? raw : 0u;   // corrupt index guard
        if (ROLE == FZ_KEY) return x.slot_base + i;
        if (x.ndict == 0) return 0u;
        if (ROLE == FZ_PRED) return (__ldg(x.pass + (i >> 5)) >> (i & 31)) & 1u;
        if (x.ddata_aligned) return __ldg((const uint32_t*)x.ddata + i);
        return ld_u32_unaligned(x.ddata + (int64_t)i * 4);
    }
    if (ROLE == FZ_PRED) {
        const int64_t v = (int64_t)(int32_t)raw;
        return (v >= x.lo && v <= x.hi) ? 1u : 0u;
    }
    if (ROLE == FZ_KEY) {
        int64_t s = (int64_t)(int32_t)raw - x.kmin;
        if ((uint64_t)s >= (uint64_t)x.range) {   // the column statistics did not cover this value
            *x.oor = 1;
            s = x.range;
        }
        return 0x80000000u | (uint32_t)s;
    }
    return raw;
}
// value k of a run -> rank pos + k.  Predicate columns: the ballot of 32 consecutive values is ORed into the rank-space
// bitmap at bit pos + k0 (two words when it straddles); lane 0 is the only writer of the bitmap.
// pass bits of a predicate column are appended, in value order, to the tile's rank-space bitmap: the writer state is the same in
// every lane (the ballots are warp-uniform), lane 0 stores each completed word
struct FzBits {
    uint64_t acc = 0;
    int fill = 0, w = 0;
};
template <int ROLE>
__device__ __forceinline__ void fz_store(uint32_t* dst, uint32_t* d, unsigned lane, bool act, uint32_t o, FzBits& bits, int n /* values of this group */) {
    // value planes: d = &plane[fz_pi(rank)], advanced by the caller (33 words per 32 ranks)
    if (ROLE == FZ_PRED) {
        const uint32_t word = __ballot_sync(FULL_MASK, act && o != 0);
        bits.acc |= (uint64_t)word << bits.fill;
        bits.fill += n;
        if (bits.fill >= 32) {
            if (lane == 0) dst[bits.w] = (uint32_t)bits.acc;
            bits.w++;
            bits.acc >>= 32;
            bits.fill -= 32;
        }
    } else if (act) {
        *d = o;
    }
}
// t values of a bit-packed run starting at value `first` of the packed area -> ranks [pos, pos + t)
template <int ROLE>
__device__ __forceinline__ void fz_unpack_bits(const FzX& x, uint32_t* dst, int pos, int t, unsigned lane, const uint8_t* bp_base, int first, int bw, FzBits& bits) {
    // lane L unpacks values L, L + 32, ...: 32 values are exactly `bw` 32-bit words, so the word pointer advances by bw per
    // step and the sub-word shift is a per-lane constant of the run
    const int64_t bit0 = (int64_t)(first + (int)lane) * bw;
    const uintptr_t qa = (uintptr_t)(bp_base + (bit0 >> 3));
    const uint32_t* wp = (const uint32_t*)(qa & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(qa & 3) * 8 + (unsigned)(bit0 & 7);
    const uint32_t vmask = bw >= 32 ? 0xffffffffu : ((1u << bw) - 1u);
    uint32_t* d = dst + fz_pi(pos + (int)lane);
    int k0 = 0;
    for (; k0 + 256 <= t; k0 += 256) {   // 8 groups per step: all sixteen loads are issued before the first value is used
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            lo[u] = __ldg(wp + u * bw);
            hi[u] = __ldg(wp + u * bw + 1);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) fz_store<ROLE>(dst, d + 33 * u, lane, true, fz_xform<ROLE>(x, __funnelshift_r(lo[u], hi[u], sh) & vmask), bits, 32);
        wp += 8 * bw;
        d += 8 * 33;
    }
    for (; k0 + 128 <= t; k0 += 128) {   // 4 groups
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            lo[u] = __ldg(wp + u * bw);
            hi[u] = __ldg(wp + u * bw + 1);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) fz_store<ROLE>(dst, d + 33 * u, lane, true, fz_xform<ROLE>(x, __funnelshift_r(lo[u], hi[u], sh) & vmask), bits, 32);
        wp += 4 * bw;
        d += 4 * 33;
    }
    for (; k0 < t; k0 += 32) {
        const bool act = k0 + (int)lane < t;
        uint32_t o = 0;
        if (act) o = fz_xform<ROLE>(x, __funnelshift_r(__ldg(wp), __ldg(wp + 1), sh) & vmask);
        fz_store<ROLE>(dst, d, lane, act, o, bits, min(32, t - k0));
        wp += bw;
        d += 33;
    }
}
template <int ROLE>
__device__ __forceinline__ void fz_unpack_const(const FzX& x, uint32_t* dst, int pos, int t, unsigned lane, uint32_t raw, FzBits& bits) {
    const uint32_t o = fz_xform<ROLE>(x, raw);
    uint32_t* d = dst + fz_pi(pos + (int)lane);
    for (int k0 = 0; k0 < t; k0 += 32, d += 33) fz_store<ROLE>(dst, d, lane, k0 + (int)lane < t, o, bits, min(32, t - k0));
}
template <int ROLE>
__device__ __forceinline__ void fz_unpack_plain(const FzX& x, uint32_t* dst, int pos, int t, unsigned lane, const uint8_t* first, FzBits& bits) {
    const uintptr_t ba = (uintptr_t)first;   // no alignment guarantee: page payloads sit at arbitrary file offsets
    const uint32_t* wp = (const uint32_t*)(ba & ~(uintptr_t)3) + lane;
    const unsigned bsh = (unsigned)(ba & 3) * 8;
    uint32_t* d = dst + fz_pi(pos + (int)lane);
    for (int k0 = 0; k0 < t; k0 += 32, wp += 32, d += 33) {
        const bool act = k0 + (int)lane < t;
        uint32_t o = 0;
        if (act) o = fz_xform<ROLE>(x, bsh ? __funnelshift_r(__ldg(wp), __ldg(wp + 1), bsh) : __ldg(wp));
        fz_store<ROLE>(dst, d, lane, act, o, bits, min(32, t - k0));
    }
}
// validity word + rank base of this lane's rows
__device__ __forceinline__ void fz_rows(const FzColumn& C, int T, int n_tile, unsigned lane, uint32_t* w_out, int* pref_out) {
    const int cnt = n_tile - 32 * (int)lane;
    uint32_t w = cnt >= 32 ? 0xffffffffu : (cnt > 0 ? (1u << cnt) - 1u : 0u);
    if (C.valid) w &= __ldg(C.valid + (int64_t)T * (FZ_TILE / 32) + lane);
    const int pc = __popc(w);
    int inc = pc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, inc, d);
        if ((int)lane >= d) inc += t;
    }
    *w_out = w;
    *pref_out = inc - pc;
}
// unpack role-column c of tile T: values / pass bits to shared memory.  Not inlined: the three roles are shared by every
// instantiation of the kernel below.
template <int ROLE>
__device__ __forceinline__ bool fz_load_col(const FzLaunch& L, int c, int T, int n_tile, uint32_t* dst, unsigned lane) {
    bool all_dict = true;
    const FzColumn& C = L.col[c];
    if (ROLE == FZ_PRED) {
        dst[lane] = 0;
        if (lane < 2) dst[32 + lane] = 0;
        __syncwarp();
    }
    FzX x;
    x.lo = C.lo;
    x.hi = C.hi;
    x.kmin = L.kmin;
    x.range = L.range;
    x.oor = L.oor;
    FzBits bits;
    int seg = __ldg(C.first_seg + T), covered = 0, pos = 0;
    for (int guard = 0; covered < n_tile && guard < FZ_TILE; guard++, seg++) {
        const FzSeg* sp = C.segs + seg;
        const int4 s0 = __ldg((const int4*)sp);   // page, row0, n, nvalid
        const int nvs = min(s0.w, FZ_TILE - pos);
        covered += s0.z > 0 ? s0.z : FZ_TILE;
        if (nvs <= 0) continue;
        const int4 s1 = __ldg((const int4*)sp + 1);   // v0 (2 words), idx.p_off, idx.run_remaining
        const int4 s2 = __ldg((const int4*)sp + 2);   // idx.bp_base_off, idx.bp_consumed, idx.rle_value, idx.is_rle
        const int4 s3 = __ldg((const int4*)sp + 3);   // vals (2 words), ddata (2 words)
        const int4 s4 = __ldg((const int4*)sp + 4);   // val_len, bw, ndict, dict_id
        const uint8_t* vals = (const uint8_t*)(((uint64_t)(uint32_t)s3.y << 32) | (uint32_t)s3.x);
        const int bw = s4.y;
        x.dict = bw >= 0;
        all_dict = all_dict && x.dict;
        x.ndict = (uint32_t)s4.z;
        x.ddata = (const uint8_t*)(((uint64_t)(uint32_t)s3.w << 32) | (uint32_t)s3.z);
        x.ddata_aligned = ((uintptr_t)x.ddata & 3) == 0;
        x.pass = nullptr;
        x.slot_base = 0;
        if (x.dict) {
            if (ROLE == FZ_PRED) x.pass = C.pass_bits + __ldg(C.pass_off + s4.w);
            if (ROLE == FZ_KEY) x.slot_base = (uint32_t)__ldg(C.dslot_base + s4.w);
            Hybrid idx;
            hybrid_restore(idx, HybridCk{s1.z, s1.w, s2.x, s2.y, (uint32_t)s2.z, s2.w}, vals + 1, vals + s4.x, bw);
            int done = 0;
            while (done < nvs) {
                if (idx.run_remaining == 0) idx.next_run();
                const int t = min(nvs - done, idx.run_remaining);
                if (idx.is_rle) fz_unpack_const<ROLE>(x, dst, pos + done, t, lane, idx.rle_value, bits);
                else fz_unpack_bits<ROLE>(x, dst, pos + done, t, lane, idx.bp_base, idx.bp_consumed, bw, bits);
                done += t;
                idx.run_remaining -= t;
                if (!idx.is_rle) idx.bp_consumed += t;
            }
        } else {   // PLAIN INT32: value k of the segment is the 32-bit word at vals + 4 (v0 + k)
            const int64_t v0 = (int64_t)(((uint64_t)(uint32_t)s1.y << 32) | (uint32_t)s1.x);
            fz_unpack_plain<ROLE>(x, dst, pos, nvs, lane, vals + v0 * 4, bits);
        }
        pos += nvs;
    }
    if (ROLE == FZ_PRED && bits.fill > 0 && lane == 0) dst[bits.w] = (uint32_t)bits.acc;   // the last, partial word
    __syncwarp();
    return all_dict;
}
// ask L2 for the page bytes of tile T of every column (lane c looks up column c, then all lanes issue the line prefetches)
__device__ __forceinline__ void fz_prefetch_tile(const FzLaunch& L, int T, unsigned lane) {
    const uint8_t* beg = nullptr;
    int bytes = 0;
    if ((int)lane < L.ncols) {
        const FzColumn& C = L.col[lane];
        const FzSeg* sp = C.segs + __ldg(C.first_seg + T);
        const int4 s0 = __ldg((const int4*)sp);
        const int4 s1 = __ldg((const int4*)sp + 1);
        const int4 s2 = __ldg((const int4*)sp + 2);
        const int4 s3 = __ldg((const int4*)sp + 3);
        const int4 s4 = __ldg((const int4*)sp + 4);
        const uint8_t* vals = (const uint8_t*)(((uint64_t)(uint32_t)s3.y << 32) | (uint32_t)s3.x);
        if (s4.y >= 0) {   // index stream from the checkpoint on (later segments of the tile start a new page: not prefetched)
            const int64_t off = s2.w ? (int64_t)s1.z : (int64_t)s2.x + (((int64_t)s2.y * s4.y) >> 3);
            beg = vals + 1 + off;
            bytes = (int)min((int64_t)((s0.w * s4.y) >> 3) + 64, (int64_t)s4.x - off);
        } else {
            const int64_t v0 = (int64_t)(((uint64_t)(uint32_t)s1.y << 32) | (uint32_t)s1.x);
            beg = vals + v0 * 4;
            bytes = s0.w * 4;
        }
    }
    for (int c = 0; c < L.ncols; c++) {
        const uint8_t* b = (const uint8_t*)__shfl_sync(FULL_MASK, (unsigned long long)beg, c);
        const int n = __shfl_sync(FULL_MASK, bytes, c);
        for (int o = 128 * (int)lane; o < n; o += 128 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
    }
}
// bits[0, popc(w)) deposited into the set positions of w (PDEP)
__device__ __forceinline__ uint32_t fz_deposit(uint32_t bits, uint32_t w) {
    if (w == 0xffffffffu) return bits;
    uint32_t out = 0;
    while (w) {
        const uint32_t b = w & (0u - w);
        if (bits & 1u) out |= b;
        bits >>= 1;
        w ^= b;
    }
    return out;
}
// The selected rows of this lane -> accumulators.  Everything that does not depend on the row is decided before the loop;
// ALLDICT (every key of the tile came from a dictionary page and none is NULL, the common case) drops the per-row choice between
// the dictionary-space and the direct arrays.  SIG != 0: the accumulator list itself is a compile-time constant (6 bits per
// accumulator: operation, argument plane, "has a valid-flag array") -- the loop body is then straight-line code; the shapes that
// TPC-DS aggregates produce most are instantiated, any other list runs the SIG = 0 body that reads the list at run time.
constexpr uint32_t fz_sig1(int op, int plane /* -1 COUNT(*), 0 key, 1 + v */, bool flag) { return 0x20u | (uint32_t)op | ((uint32_t)(plane + 1) << 2) | (flag ? 0x10u : 0u); }
constexpr uint32_t FZ_SIG_SUM_COUNT = fz_sig1(0, 1, false) | (fz_sig1(1, 1, false) << 6);          // SUM(x), COUNT(x)   (AVG's partial state)
constexpr uint32_t FZ_SIG_SUM = fz_sig1(0, 1, true);                                               // SUM(x)
constexpr uint32_t FZ_SIG_COUNT_STAR = fz_sig1(1, -1, false);                                      // COUNT(*)
constexpr uint32_t FZ_SIG_SUM_COUNT_STAR = fz_sig1(0, 1, true) | (fz_sig1(1, -1, false) << 6);     // SUM(x), COUNT(*)
constexpr uint32_t FZ_SIG_SUM_SUM = fz_sig1(0, 1, true) | (fz_sig1(0, 2, true) << 6);              // SUM(x), SUM(y)
// key_at(rank) -> the key's accumulator slot word (bit 31: direct table), val_at(v, rank) -> argument v's value
template <int NV, int NACC, bool ALLDICT, uint32_t SIG, class KF, class VF>
__device__ __forceinline__ void fz_rows_to_accs(const FzLaunch& L, uint32_t sel, uint32_t wk, int prefk, const uint32_t* wv, const int* prefv,
                                                KF key_at, VF val_at) {
    constexpr int NVR = NV < 0 ? FZ_MAX_COLS - 1 : (NV == 0 ? 1 : NV);
    constexpr int NA = NACC < 0 ? FZ_MAX_ACCS : NACC;
    const int nacc = NACC < 0 ? L.nacc : NACC;
    unsigned long long* bd[NA];   // dictionary-space base
    unsigned long long* bx[NA];   // direct base
    uint8_t* vd[NA];
    uint8_t* vx[NA];
    int op[NA], plane[NA];        // op: 0 add value, 1 add one, 2 min, 3 max ; plane: -1 COUNT(*), 0 key validity, 1 + v argument plane v
    bool flag[NA];
    bool need_seen = true;
#pragma unroll
    for (int a = 0; a < NA; a++) {
        bd[a] = bx[a] = nullptr;
        vd[a] = vx[a] = nullptr;
        op[a] = 0;
        plane[a] = -1;
        flag[a] = false;
        if (a < nacc) {
            const FzAcc& A = L.acc[a];
            bd[a] = A.dspace;
            bx[a] = A.direct;
            vd[a] = A.dspace_valid;
            vx[a] = A.direct_valid;
            if (SIG) {
                constexpr uint32_t dummy = 0;
                (void)dummy;
                const uint32_t s6 = (SIG >> (6 * a)) & 0x3fu;
                op[a] = (int)(s6 & 3u);
                plane[a] = (int)((s6 >> 2) & 3u) - 1;
                flag[a] = (s6 & 0x10u) != 0;
            } else {
                op[a] = A.kind == ACC_COUNT ? 1 : A.kind == ACC_MIN ? 2 : A.kind == ACC_MAX ? 3 : 0;
                plane[a] = A.col < 0 ? -1 : A.col - L.npred;
                flag[a] = A.direct_valid != nullptr;
            }
            if (op[a] == 1 && plane[a] < 0) need_seen = false;   // COUNT(*) marks every selected row's group
        }
    }
    const uint32_t null_slot = 0x80000000u | (uint32_t)L.range;
    // Software pipeline: the key and the arguments of the next selected row are fetched (shared-memory windows, dictionary
    // lookups) before the atomics of the current row are issued, so the two latencies overlap.
    struct Row {
        int i;
        uint32_t U;
        long long val[NVR];
        bool ok[NVR];
    };
    auto fetch = [&](Row& R) {
        R.i = __ffs(sel) - 1;
        sel &= sel - 1;
        const uint32_t below = (1u << R.i) - 1u;
        R.U = null_slot;
        if (ALLDICT || ((wk >> R.i) & 1u)) R.U = key_at(prefk + __popc(wk & below));
#pragma unroll
        for (int v = 0; v < NVR; v++) {
            R.ok[v] = (wv[v] >> R.i) & 1u;
            R.val[v] = R.ok[v] ? (long long)(int32_t)val_at(v, prefv[v] + __popc(wv[v] & below)) : 0ll;
        }
    };
    auto commit = [&](const Row& R) {
        const bool dsp = ALLDICT ? true : !(R.U >> 31);
        const int64_t slot = (int64_t)(R.U & 0x7fffffffu);
        if (SIG == FZ_SIG_SUM_COUNT && L.pack_shift > 0 && dsp) {   // both accumulators of the row in one atomic
            if (R.ok[0]) {
                const unsigned long long d = (unsigned long long)(R.val[0] - L.pack_bias);
                if (d >> L.pack_bits) *L.oor = 1;   // the column statistics did not cover this value
                else atomicAdd(bd[0] + slot, (1ull << L.pack_shift) | d);
            } else {
                L.seen_dspace[slot] = 1;
            }
            return;
        }
        bool marked = false;
#pragma unroll
        for (int a = 0; a < NA; a++) {
            if (a >= nacc) break;
            long long v = 1;
            bool valid = true;
            if (plane[a] == 0) valid = ALLDICT ? true : ((wk >> R.i) & 1u);
#pragma unroll
            for (int q = 0; q < NVR; q++)
                if (plane[a] == 1 + q) {
                    valid = R.ok[q];
                    v = R.val[q];
                }
            if (!valid) continue;
            unsigned long long* p = (dsp ? bd[a] : bx[a]) + slot;
            if (op[a] == 1) atomicAdd(p, 1ull);
            else if (op[a] == 0) atomicAdd(p, (unsigned long long)v);   // SUM (wrapping, sum.rs:115)
            else if (op[a] == 2) atomicMin((long long*)p, v);
            else atomicMax((long long*)p, v);
            if (flag[a]) {
                (dsp ? vd[a] : vx[a])[slot] = 1;
                marked = true;
            }
            marked = marked || op[a] == 1;
        }
        if (need_seen && !marked) (dsp ? L.seen_dspace : L.seen_direct)[slot] = 1;   // the group exists although no accumulator shows it
    };
    if (!sel) return;
    Row cur, nxt;
    fetch(cur);
    while (sel) {
        fetch(nxt);
        commit(cur);
        cur = nxt;
    }
    commit(cur);
}
// ------------------------------------------------------------------------------------------------ TMA-staged tiles
// Tiles whose columns are each ONE segment of a "regular" index stream (or of a PLAIN page) need no run walk: value k of the page
// sits at an arithmetic bit position.  For those the packed bytes of the tile -- 128 bytes per bit of width and column -- are
// brought into shared memory with one bulk copy per column (cp.async.bulk, completion on an mbarrier; SASS UBLKCP), double
// buffered: a persistent warp issues the copies of its next tile before it works on the current one, so global latency is off
// the critical path, and every access to the packed stream is a shared-memory window read.  Nothing is unpacked that is not
// used: predicate columns are decoded for the rows of the lane, key and argument columns only for the rows that pass.
// Tiles that do not qualify (a page boundary inside the tile, RLE runs, short literal runs, wide PLAIN pages in a dictionary
// column) are left to fz_kernel, which skips the tiles done here by evaluating the same plan.
struct FzStagePlan {      // lane c: column c of the tile
    const uint8_t* src;   // 16-byte aligned first byte to copy
    int32_t bytes;        // multiple of 16, 0 = the tile holds no value of this column
    int32_t sub;          // stream offset of src (relative to the first byte of the index stream / of the PLAIN values)
    int32_t bw, v0, ndict, dict_id;
    const uint8_t* ddata;
};
__device__ __forceinline__ bool fz_stage_plan(const FzLaunch& L, int T, unsigned lane, FzStagePlan& pl) {
    bool ok = true;
    pl.src = nullptr;
    pl.bytes = 0;
    pl.sub = 0;
    pl.bw = 0;
    pl.v0 = 0;
    pl.ndict = 0;
    pl.dict_id = 0;
    pl.ddata = nullptr;
    if ((int)lane < L.ncols) {
        const FzColumn& C = L.col[lane];
        const int n_tile = (int)min((int64_t)FZ_TILE, L.n_rows - (int64_t)T * FZ_TILE);
        const FzSeg* sp = C.segs + __ldg(C.first_seg + T);
        const int4 s0 = __ldg((const int4*)sp);       // page, row0, n, nvalid
        const int4 s1 = __ldg((const int4*)sp + 1);   // v0 (2 words), ...
        const int4 s3 = __ldg((const int4*)sp + 3);   // vals (2 words), ddata (2 words)
        const int4 s4 = __ldg((const int4*)sp + 4);   // val_len, bw, ndict, dict_id
        const int4 s5 = __ldg((const int4*)sp + 5);   // flags
        const int64_t v0 = (int64_t)(((uint64_t)(uint32_t)s1.y << 32) | (uint32_t)s1.x);
        const uint8_t* vals = (const uint8_t*)(((uint64_t)(uint32_t)s3.y << 32) | (uint32_t)s3.x);
        const int bw = s4.y, nv = s0.w;
        ok = C.stage_cap > 0 && s0.z == n_tile && (s5.x & 1) && (bw < 0 || (bw >= 1 && bw <= 32)) && v0 + nv < (int64_t)0x7fffffff;
        if (ok && nv > 0) {
            int64_t fb, lb;   // stream bytes [fb, lb) hold the tile's values
            const uint8_t* stream;
            if (bw < 0) {
                stream = vals;
                fb = 4 * v0;
                lb = 4 * (v0 + nv);
            } else {
                stream = vals + 1;
                const int64_t k1 = v0 + nv - 1, q0 = v0 / 504, q1 = k1 / 504, run_bytes = 1 + 63 * bw;
                fb = q0 * run_bytes + 1 + (((v0 - 504 * q0) * bw) >> 3);
                lb = q1 * run_bytes + 1 + (((k1 - 504 * q1) * bw + bw + 7) >> 3);
            }
            const uintptr_t a0 = (uintptr_t)(stream + fb) & ~(uintptr_t)15, a1 = ((uintptr_t)(stream + lb) + 15) & ~(uintptr_t)15;
            pl.src = (const uint8_t*)a0;
            pl.bytes = (int32_t)(a1 - a0);
            pl.sub = (int32_t)((int64_t)a0 - (int64_t)(uintptr_t)stream);
            ok = (int64_t)(a1 - a0) + 32 <= C.stage_cap && lb + (bw >= 0 ? 1 : 0) <= (int64_t)s4.x;   // (+32: 8-byte windows, up to three values behind the last one)
        }
        pl.bw = bw;
        pl.v0 = (int32_t)v0;
        pl.ndict = s4.z;
        pl.dict_id = s4.w;
        pl.ddata = (const uint8_t*)(((uint64_t)(uint32_t)s3.w << 32) | (uint32_t)s3.z);
    }
    return __all_sync(FULL_MASK, ok);
}
__device__ __forceinline__ uint32_t fz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fz_mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fz_smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void fz_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fz_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fz_mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fz_smem_u32(b)) : "memory"); }
__device__ __forceinline__ void fz_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fz_smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(fz_smem_u32(b))
                 : "memory");
}
// false: the phase did not complete within ~2^26 polls (a lost copy: reported, never spun on forever)
__device__ __forceinline__ bool fz_mbar_wait(uint64_t* b, uint32_t parity) {
    for (int spin = 0; spin < (1 << 26); spin++) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(fz_smem_u32(b)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
struct FzStaged {        // one column of the current tile, as every lane sees it
    uint32_t st;         // shared-window address of the column's stage
    int32_t adj;         // stream byte x sits at st + x + adj ... for run q, value r of the run: st + q * run_bytes + adj + (r * bw >> 3)
    uint32_t bw, run_bytes, hdr_bits;   // PLAIN pages are "runs" of 504 32-bit values without a header: bw 32, run_bytes 2016
    uint32_t mask;
    int32_t v0, ndict;
    bool dict;
    uint32_t aux;        // predicate: first word of the dictionary's pass bits; key: first slot of the dictionary
    const uint8_t* ddata;
};
__device__ __forceinline__ uint32_t fz_lds(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));   // (volatile: the async proxy rewrites the stage between tiles)
    return v;
}
// bits [X, X + bw) of the stage, X a bit address
__device__ __forceinline__ uint32_t fz_staged_bits(const FzStaged& S, uint32_t X) {
    const uint32_t a = S.st + ((X >> 3) & ~3u);
    return __funnelshift_r(fz_lds(a), fz_lds(a + 4), X & 31u) & S.mask;
}
// value k of the page (an index for dictionary pages, the value itself for PLAIN pages)
__device__ __forceinline__ uint32_t fz_staged_get(const FzStaged& S, int k) {
    const uint32_t q = (uint32_t)k / 504u, r = (uint32_t)k - q * 504u;
    return fz_staged_bits(S, 8u * (q * S.run_bytes + (uint32_t)S.adj) + r * S.bw);
}
template <int NV, int NACC, uint32_t SIG>
__global__ void __launch_bounds__(FZ_WARPS * 32) fz_staged_kernel(const __grid_constant__ FzLaunch L, int per_warp_bytes, int nstage) {
    extern __shared__ __align__(16) uint8_t fz_smem[];
    const int wid = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    uint8_t* base = fz_smem + (size_t)wid * per_warp_bytes;
    uint64_t* mbar = (uint64_t*)base;              // [2]
    int32_t* meta = (int32_t*)(base + 16);         // [2][FZ_MAX_COLS][8]
    uint8_t* stage0 = base + 16 + 2 * FZ_MAX_COLS * 32;
    int my_off = 0, stage_bytes = 0;               // lane c: offset of column c inside a stage
    for (int c = 0; c < L.ncols; c++) {
        if (c == (int)lane) my_off = stage_bytes;
        stage_bytes += L.col[c].stage_cap;
    }
    if (lane == 0) {
        fz_mbar_init(&mbar[0], L.ncols);
        fz_mbar_init(&mbar[1], L.ncols);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int T, int s) -> bool {
        FzStagePlan pl;
        if (!fz_stage_plan(L, T, lane, pl)) return false;
        if ((int)lane < L.ncols) {
            const FzColumn& C = L.col[lane];
            int32_t* m = meta + (s * FZ_MAX_COLS + (int)lane) * 8;
            uint32_t aux = 0;
            if (pl.bw >= 0 && C.role == FZ_PRED) aux = (uint32_t)__ldg(C.pass_off + pl.dict_id);
            if (pl.bw >= 0 && C.role == FZ_KEY) aux = (uint32_t)__ldg(C.dslot_base + pl.dict_id);
            m[0] = pl.sub;
            m[1] = pl.bw;
            m[2] = pl.v0;
            m[3] = pl.ndict;
            m[4] = (int32_t)aux;
            m[5] = (int32_t)(uint32_t)(uintptr_t)pl.ddata;
            m[6] = (int32_t)(uint32_t)((uintptr_t)pl.ddata >> 32);
            if (pl.bytes > 0) {
                fz_mbar_expect_tx(&mbar[s], (uint32_t)pl.bytes);
                fz_bulk_load(stage0 + (size_t)s * stage_bytes + my_off, pl.src, (uint32_t)pl.bytes, &mbar[s]);
            } else {
                fz_mbar_arrive(&mbar[s]);
            }
        }
        return true;
    };
    auto column = [&](int c, int s) -> FzStaged {
        const int32_t* m = meta + (s * FZ_MAX_COLS + c) * 8;
        FzStaged S;
        int off = 0;
        for (int q = 0; q < c; q++) off += L.col[q].stage_cap;
        S.st = fz_smem_u32(stage0 + (size_t)s * stage_bytes + off);
        const int bw = m[1];
        S.dict = bw >= 0;
        S.bw = S.dict ? (uint32_t)bw : 32u;
        S.run_bytes = S.dict ? 1u + 63u * (uint32_t)bw : 2016u;
        S.hdr_bits = S.dict ? 8u : 0u;
        S.adj = (S.dict ? 1 : 0) - m[0];
        S.mask = S.bw >= 32u ? 0xffffffffu : (1u << S.bw) - 1u;
        S.v0 = m[2];
        S.ndict = m[3];
        S.aux = (uint32_t)m[4];
        S.ddata = (const uint8_t*)(((uint64_t)(uint32_t)m[6] << 32) | (uint32_t)m[5]);
        return S;
    };
    uint32_t phase = 0;
    int n_done = 0;
    // Tiles are handed out by a counter: the kernel shares the GPU with the decompression and scout kernels of the next batch, so
    // only part of the grid is resident at first -- a fixed tile-to-warp map would leave the late blocks' share for the end.
    auto grab = [&]() -> int {
        int t = 0;
        if (lane == 0) t = atomicAdd(L.left + 1, 1);
        return __shfl_sync(FULL_MASK, t, 0);
    };
    int T = grab(), s = 0;
    const int toggle = nstage == 2 ? 1 : 0;   // one stage: the copies of a tile are issued when the tile before it is done (more warps per SM instead)
    bool staged = toggle && T < L.n_tiles ? issue(T, 0) : false;
    for (int Tn = 0; T < L.n_tiles; T = Tn, s ^= toggle) {
        __syncwarp();   // every lane is done with the other stage (the tile before this one)
        Tn = grab();
        bool mine;
        if (toggle) {
            const bool staged_next = Tn < L.n_tiles ? issue(Tn, s ^ 1) : false;
            mine = staged;
            staged = staged_next;
        } else {
            mine = issue(T, 0);
        }
        if (!mine) {   // left to the tile kernel
            if (lane == 0) L.left[2 + atomicAdd(L.left, 1)] = T;
            continue;
        }
        if (!fz_mbar_wait(&mbar[s], (phase >> s) & 1u)) {
            *L.oor = 2;   // (the operator restarts the scan unfused)
            return;
        }
        phase ^= 1u << s;
        n_done++;
        const int n_tile = (int)min((int64_t)FZ_TILE, L.n_rows - (int64_t)T * FZ_TILE);
        const int cnt = n_tile - 32 * (int)lane;
        const uint32_t rowmask = cnt >= 32 ? 0xffffffffu : (cnt > 0 ? (1u << cnt) - 1u : 0u);
        // ---- 1. predicates: the pass bit of every value of this lane's rows
        uint32_t sel = rowmask;
        for (int p = 0; p < L.npred; p++) {
            uint32_t w;
            int pref;
            fz_rows(L.col[p], T, n_tile, lane, &w, &pref);
            const FzStaged S = column(p, s);
            const FzColumn& C = L.col[p];
            const uint32_t* pass = C.pass_bits + S.aux;
            uint32_t res = 0;
            const int k0 = S.v0 + pref;
            if (p == 0) {   // every valid row is looked at: walk the lane's values in stream order (no division per value)
                uint32_t q = (uint32_t)k0 / 504u, r = (uint32_t)k0 - q * 504u;
                uint32_t X = 8u * (q * S.run_bytes + (uint32_t)S.adj) + r * S.bw;
                // four values per round: the eight window reads and the four pass-bit loads are issued before the first is used
                // (reads behind the lane's last value stay inside the stage: it carries 32 spare bytes)
                uint32_t ww = w;
                for (int n = __popc(w); n > 0; n -= 4) {
                    uint32_t raw[4], bit[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        raw[u] = fz_staged_bits(S, X);
                        X += S.bw;
                        if (++r == 504u) {
                            r = 0;
                            X += S.hdr_bits;
                        }
                    }
                    if (S.dict) {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t e = raw[u] < (uint32_t)S.ndict ? raw[u] : 0u;
                            bit[u] = S.ndict > 0 ? (__ldg(pass + (e >> 5)) >> (e & 31)) & 1u : 0u;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int64_t v = (int64_t)(int32_t)raw[u];
                            bit[u] = (v >= C.lo && v <= C.hi) ? 1u : 0u;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++)
                        if (u < n) {
                            const int i = __ffs(ww) - 1;
                            ww &= ww - 1;
                            res |= bit[u] << i;
                        }
                }
            } else {        // rows already dropped by an earlier predicate are not looked at
                for (uint32_t ww = w & sel; ww;) {
                    const int i = __ffs(ww) - 1;
                    ww &= ww - 1;
                    const uint32_t raw = fz_staged_get(S, k0 + __popc(w & ((1u << i) - 1u)));
                    uint32_t bit;
                    if (S.dict) {
                        const uint32_t e = raw < (uint32_t)S.ndict ? raw : 0u;
                        bit = S.ndict > 0 ? (__ldg(pass + (e >> 5)) >> (e & 31)) & 1u : 0u;
                    } else {
                        const int64_t v = (int64_t)(int32_t)raw;
                        bit = (v >= C.lo && v <= C.hi) ? 1u : 0u;
                    }
                    res |= bit << i;
                }
            }
            sel &= res;
        }
        int nsel = __popc(sel);
#pragma unroll
        for (int d = 16; d; d >>= 1) nsel += __shfl_xor_sync(FULL_MASK, nsel, d);
        if (nsel == 0) continue;
        if (lane == 0) atomicAdd(L.selected_rows, (unsigned long long)nsel);
        // ---- 2. selected rows -> accumulators; keys and arguments are unpacked on demand
        uint32_t wk;
        int prefk;
        fz_rows(L.col[L.key_col], T, n_tile, lane, &wk, &prefk);
        const FzStaged K = column(L.key_col, s);
        constexpr int NVR = NV < 0 ? FZ_MAX_COLS - 1 : (NV == 0 ? 1 : NV);
        const int nplanes = L.ncols - L.npred;
        uint32_t wv[NVR];
        int prefv[NVR];
        FzStaged V[NVR];
#pragma unroll
        for (int v = 0; v < NVR; v++) {
            wv[v] = 0;
            prefv[v] = 0;
            V[v] = K;
            if (v < (NV < 0 ? nplanes - 1 : NV)) {
                fz_rows(L.col[L.key_col + 1 + v], T, n_tile, lane, &wv[v], &prefv[v]);
                V[v] = column(L.key_col + 1 + v, s);
            }
        }
        auto key_at = [&](int rank) -> uint32_t {
            const uint32_t raw = fz_staged_get(K, K.v0 + rank);
            if (K.dict) return K.aux + (raw < (uint32_t)K.ndict ? raw : 0u);
            int64_t slot = (int64_t)(int32_t)raw - L.kmin;
            if ((uint64_t)slot >= (uint64_t)L.range) {   // the column statistics did not cover this value
                *L.oor = 1;
                slot = L.range;
            }
            return 0x80000000u | (uint32_t)slot;
        };
        auto val_at = [&](int v, int rank) -> uint32_t {
            const FzStaged& S = V[v];
            const uint32_t raw = fz_staged_get(S, S.v0 + rank);
            if (!S.dict) return raw;
            if (S.ndict <= 0) return 0u;
            const uint32_t e = raw < (uint32_t)S.ndict ? raw : 0u;
            return ((uintptr_t)S.ddata & 3) == 0 ? __ldg((const uint32_t*)S.ddata + e) : ld_u32_unaligned(S.ddata + (int64_t)e * 4);
        };
        const bool dict_no_null = K.dict && __all_sync(FULL_MASK, wk == rowmask);   // NULL keys live in the direct table
        if (dict_no_null) fz_rows_to_accs<NV, NACC, true, SIG>(L, sel, wk, prefk, wv, prefv, key_at, val_at);
        else fz_rows_to_accs<NV, NACC, false, SIG>(L, sel, wk, prefk, wv, prefv, key_at, val_at);
    }
    if (lane == 0 && n_done) atomicAdd(L.selected_rows + 1, (unsigned long long)n_done);
}
// NV argument planes, NACC accumulators (compile time: the row loop is fully unrolled, descriptors come straight from the
// constant bank); NV = -1: any shape, loops at run time
template <int NV, int NACC, uint32_t SIG>
__device__ __forceinline__ void fz_tile(const FzLaunch& L, int T, uint8_t* fz_smem, int wid, unsigned lane) {
    fz_prefetch_tile(L, T, lane);
    const int nplanes = L.ncols - L.npred;   // key + argument columns
    // per warp: [36 words pass bits][nplanes value planes]
    const size_t per_warp = 36 * 4 + (size_t)nplanes * FZ_VSTRIDE * 4;
    uint32_t* s_bits = (uint32_t*)(fz_smem + (size_t)wid * per_warp);
    uint32_t* s_plane = s_bits + 36;
    const int n_tile = (int)min((int64_t)FZ_TILE, L.n_rows - (int64_t)T * FZ_TILE);
    // ---- 1. selection word of this lane's rows
    const int cnt = n_tile - 32 * (int)lane;
    uint32_t sel = cnt >= 32 ? 0xffffffffu : (cnt > 0 ? (1u << cnt) - 1u : 0u);
    for (int p = 0; p < L.npred; p++) {
        uint32_t w;
        int pref;
        fz_rows(L.col[p], T, n_tile, lane, &w, &pref);
        fz_load_col<FZ_PRED>(L, p, T, n_tile, s_bits, lane);
        const uint32_t bits = __funnelshift_r(s_bits[pref >> 5], s_bits[(pref >> 5) + 1], pref & 31);   // pass bits of my values, from bit 0
        sel &= fz_deposit(bits, w);
        __syncwarp();
    }
    int nsel = __popc(sel);
#pragma unroll
    for (int d = 16; d; d >>= 1) nsel += __shfl_xor_sync(FULL_MASK, nsel, d);
    if (nsel == 0) return;   // nothing of this tile survives the filter: the other columns are not even unpacked
    if (lane == 0) atomicAdd(L.selected_rows, (unsigned long long)nsel);
    // ---- 2. key and argument columns
    uint32_t wk;
    int prefk;
    fz_rows(L.col[L.key_col], T, n_tile, lane, &wk, &prefk);
    const bool key_all_dict = fz_load_col<FZ_KEY>(L, L.key_col, T, n_tile, s_plane, lane);
    constexpr int NVR = NV < 0 ? FZ_MAX_COLS - 1 : (NV == 0 ? 1 : NV);
    uint32_t wv[NVR];
    int prefv[NVR];
#pragma unroll
    for (int v = 0; v < NVR; v++) {
        wv[v] = 0;
        prefv[v] = 0;
        if (v < (NV < 0 ? nplanes - 1 : NV)) {
            fz_rows(L.col[L.key_col + 1 + v], T, n_tile, lane, &wv[v], &prefv[v]);
            fz_load_col<FZ_VALUE>(L, L.key_col + 1 + v, T, n_tile, s_plane + (size_t)(1 + v) * FZ_VSTRIDE, lane);
        }
    }
    __syncwarp();
    // ---- 3. selected rows -> accumulators
    const uint32_t rowmask = cnt >= 32 ? 0xffffffffu : (cnt > 0 ? (1u << cnt) - 1u : 0u);
    const bool dict_no_null = key_all_dict && __all_sync(FULL_MASK, wk == rowmask);   // NULL keys live in the direct table
    auto key_at = [&](int rank) -> uint32_t { return s_plane[fz_pi(rank)]; };
    auto val_at = [&](int v, int rank) -> uint32_t { return s_plane[(size_t)(1 + v) * FZ_VSTRIDE + fz_pi(rank)]; };
    if (dict_no_null) fz_rows_to_accs<NV, NACC, true, SIG>(L, sel, wk, prefk, wv, prefv, key_at, val_at);
    else fz_rows_to_accs<NV, NACC, false, SIG>(L, sel, wk, prefk, wv, prefv, key_at, val_at);
}
// one warp per tile; behind the staged kernel (L.staged): persistent warps over the tiles it left (L.left: count, then the tiles)
template <int NV, int NACC, uint32_t SIG>
__global__ void __launch_bounds__(FZ_WARPS * 32) fz_kernel(const __grid_constant__ FzLaunch L) {
    extern __shared__ __align__(16) uint8_t fz_smem[];
    const int wid = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (!L.staged) {
        const int T = blockIdx.x * FZ_WARPS + wid;
        if (T < L.n_tiles) fz_tile<NV, NACC, SIG>(L, T, fz_smem, wid, lane);
        return;
    }
    const int count = __ldg(L.left);
    for (int i = blockIdx.x * FZ_WARPS + wid; i < count; i += gridDim.x * FZ_WARPS) {
        fz_tile<NV, NACC, SIG>(L, __ldg(L.left + 2 + i), fz_smem, wid, lane);
        __syncwarp();
    }
}
static size_t fz_smem_bytes(int nplanes) { return (size_t)FZ_WARPS * (36 * 4 + (size_t)nplanes * FZ_VSTRIDE * 4); }
template <int NV, int NACC, uint32_t SIG>
static void fz_launch(Ctx& ctx, const FzLaunch& L, size_t smem) {
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_OK(cudaFuncSetAttribute(fz_kernel<NV, NACC, SIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fz_smem_bytes(FZ_MAX_COLS)));
        CUDA_OK(cudaFuncSetAttribute(fz_staged_kernel<NV, NACC, SIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        attr_set = true;
    }
    if (L.staged) {
        int stage_bytes = 0;
        for (int c = 0; c < L.ncols; c++) stage_bytes += L.col[c].stage_cap;
        // One stage per warp by default (AURON_FUSED_STAGES=2: double buffering).  Measured on B200, SF100 bench, three batches in
        // flight: 5.59 ms per step with one stage, 6.02 ms with two -- half the shared memory lets twice as many warps hide the copy
        // latency themselves and leaves room for the decompression / scout kernels of the next batch on the same SMs.
        static const int nstage = getenv("AURON_FUSED_STAGES") ? std::max(1, std::min(2, atoi(getenv("AURON_FUSED_STAGES")))) : 1;
        const int per_warp = 16 + 2 * FZ_MAX_COLS * 32 + nstage * stage_bytes;
        const size_t smem2 = (size_t)FZ_WARPS * per_warp;
        int per_sm = 0;
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fz_staged_kernel<NV, NACC, SIG>, FZ_WARPS * 32, smem2));
        AURON_CHECK(per_sm > 0, "the staged scan kernel does not fit an SM");
        static int n_sm = 0;
        if (!n_sm) CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx.device));
        const int blocks = std::min((L.n_tiles + FZ_WARPS - 1) / FZ_WARPS, n_sm * per_sm);   // persistent warps, round robin over the tiles
        fz_staged_kernel<NV, NACC, SIG><<<blocks, FZ_WARPS * 32, smem2, ctx.stream>>>(L, per_warp, nstage);
        CUDA_OK(cudaGetLastError());
        launch_count(ctx);
        int per_sm1 = 0;
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, fz_kernel<NV, NACC, SIG>, FZ_WARPS * 32, smem));
        fz_kernel<NV, NACC, SIG><<<std::min((L.n_tiles + FZ_WARPS - 1) / FZ_WARPS, n_sm * std::max(per_sm1, 1)), FZ_WARPS * 32, smem, ctx.stream>>>(L);
        return;
    }
    fz_kernel<NV, NACC, SIG><<<(L.n_tiles + FZ_WARPS - 1) / FZ_WARPS, FZ_WARPS * 32, smem, ctx.stream>>>(L);
}
void fz_run(Ctx& ctx, const FzLaunch& L0) {
    if (L0.n_tiles <= 0) return;
    FzLaunch L = L0;
    {   // the staged kernel needs a stage of every column in shared memory, twice, for at least 8 warps per SM
        int stage_bytes = 0;
        bool all = true;
        for (int c = 0; c < L.ncols; c++) {
            stage_bytes += L.col[c].stage_cap;
            all = all && L.col[c].stage_cap > 0;
        }
        const bool off = getenv("AURON_FUSED_NO_TMA") != nullptr;
        L.staged = all && !off && (16 + 2 * FZ_MAX_COLS * 32 + 2 * stage_bytes) * 8 <= (200 << 10) ? 1 : 0;
    }
    Buf left;   // work list of the staged kernel and what it leaves to the tile kernel
    if (L.staged) {
        left = dalloc(ctx, ((size_t)L.n_tiles + 2) * 4);   // [0] tiles left to the tile kernel, [1] next tile to hand out, then the tiles left
        CUDA_OK(cudaMemsetAsync(left->ptr, 0, 8, ctx.stream));
        L.left = P<int32_t>(left);
    }
    const int nplanes = L.ncols - L.npred, nv = nplanes - 1;
    const size_t smem = fz_smem_bytes(nplanes);
    // signature of the accumulator list (see fz_rows_to_accs)
    uint32_t sig = 0;
    if (L.nacc <= 4 && !getenv("AURON_FUSED_GENERIC"))
        for (int a = 0; a < L.nacc; a++) {
            const FzAcc& A = L.acc[a];
            const int op = A.kind == ACC_COUNT ? 1 : A.kind == ACC_MIN ? 2 : A.kind == ACC_MAX ? 3 : 0;
            const int plane = A.col < 0 ? -1 : A.col - L.npred;
            if (plane > 2) {
                sig = 0;
                break;
            }
            sig |= fz_sig1(op, plane, A.direct_valid != nullptr) << (6 * a);
        }
    ProfScope ps(ctx, "fz_scan_filter_agg");
    if (sig == FZ_SIG_SUM_COUNT && nv == 1 && L.nacc == 2) fz_launch<1, 2, FZ_SIG_SUM_COUNT>(ctx, L, smem);
    else if (sig == FZ_SIG_SUM && nv == 1 && L.nacc == 1) fz_launch<1, 1, FZ_SIG_SUM>(ctx, L, smem);
    else if (sig == FZ_SIG_COUNT_STAR && nv == 0 && L.nacc == 1) fz_launch<0, 1, FZ_SIG_COUNT_STAR>(ctx, L, smem);
    else if (sig == FZ_SIG_SUM_COUNT_STAR && nv == 1 && L.nacc == 2) fz_launch<1, 2, FZ_SIG_SUM_COUNT_STAR>(ctx, L, smem);
    else if (sig == FZ_SIG_SUM_SUM && nv == 2 && L.nacc == 2) fz_launch<2, 2, FZ_SIG_SUM_SUM>(ctx, L, smem);
    else fz_launch<-1, -1, 0>(ctx, L, smem);
    LAUNCH_CHECK(ctx);
}

// ------------------------------------------------------------------------------------------------------------ dictionary space
__global__ void __launch_bounds__(256) fz_fill_u64_kernel(unsigned long long* p, int64_t n, unsigned long long v) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}
void fz_init_dspace(Ctx& ctx, const FzLaunch& L, int64_t dict_slots) {
    if (dict_slots <= 0) return;
    for (int a = 0; a < L.nacc; a++) {
        const unsigned long long init = L.acc[a].kind == ACC_MIN ? 0x7fffffffffffffffull : L.acc[a].kind == ACC_MAX ? 0x8000000000000000ull : 0ull;
        if (init == 0) CUDA_OK(cudaMemsetAsync(L.acc[a].dspace, 0, (size_t)dict_slots * 8, ctx.stream));
        else {
            fz_fill_u64_kernel<<<(unsigned)((dict_slots + 255) / 256), 256, 0, ctx.stream>>>(L.acc[a].dspace, dict_slots, init);
            LAUNCH_CHECK(ctx);
        }
        if (L.acc[a].dspace_valid) CUDA_OK(cudaMemsetAsync(L.acc[a].dspace_valid, 0, (size_t)dict_slots, ctx.stream));
    }
    CUDA_OK(cudaMemsetAsync(L.seen_dspace, 0, (size_t)dict_slots, ctx.stream));
}
// one thread per dictionary entry: whatever the batch accumulated under the entry moves to the direct table at key - kmin
__global__ void __launch_bounds__(256) fz_merge_kernel(const __grid_constant__ FzMerge M, int64_t dict_slots) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= dict_slots) return;
    bool any = M.seen_dspace[g] != 0;
    unsigned long long v[FZ_MAX_ACCS];
    unsigned long long packed_cnt = 0;
    bool vb[FZ_MAX_ACCS];
#pragma unroll
    for (int a = 0; a < FZ_MAX_ACCS; a++) {
        v[a] = 0;
        vb[a] = false;
        if (a < M.nacc) {
            v[a] = M.acc[a].dspace[g];
            if (M.pack_shift > 0 && a == 0) {   // SUM and COUNT of the entry in one word (see FzLaunch)
                packed_cnt = v[0] >> M.pack_shift;
                v[0] = (v[0] & ((1ull << M.pack_shift) - 1ull)) + packed_cnt * (unsigned long long)M.pack_bias;
            } else if (M.pack_shift > 0 && a == 1) {
                v[1] += packed_cnt;
            }
            vb[a] = M.acc[a].dspace_valid && M.acc[a].dspace_valid[g];
            const unsigned long long init = M.acc[a].kind == ACC_MIN ? 0x7fffffffffffffffull : M.acc[a].kind == ACC_MAX ? 0x8000000000000000ull : 0ull;
            any = any || v[a] != init || vb[a];
        }
    }
    if (!any) return;
    int lo = 0, hi = M.n_dicts - 1;
    while (lo < hi) {   // last dictionary whose first slot is <= g
        const int m = (lo + hi + 1) >> 1;
        if (M.dslot_base[m] <= g) lo = m;
        else hi = m - 1;
    }
    const PqDict d = M.dicts[lo];
    const int64_t i = g - M.dslot_base[lo];
    int64_t slot = (int64_t)(int32_t)ld_u32_unaligned(d.data + i * 4) - M.kmin;
    if ((uint64_t)slot >= (uint64_t)M.range) {
        *M.oor = 1;
        return;
    }
#pragma unroll
    for (int a = 0; a < FZ_MAX_ACCS; a++)
        if (a < M.nacc) {
            unsigned long long* p = M.acc[a].direct + slot;
            switch (M.acc[a].kind) {
                case ACC_MIN: if (v[a] != 0x7fffffffffffffffull) atomicMin((long long*)p, (long long)v[a]); break;
                case ACC_MAX: if (v[a] != 0x8000000000000000ull) atomicMax((long long*)p, (long long)v[a]); break;
                default: if (v[a]) atomicAdd(p, v[a]); break;
            }
            if (vb[a] && M.acc[a].direct_valid) M.acc[a].direct_valid[slot] = 1;
        }
    if (M.seen_dspace[g]) M.seen_direct[slot] = 1;
}
void fz_merge(Ctx& ctx, const FzMerge& M, int64_t dict_slots) {
    if (dict_slots <= 0 || M.n_dicts <= 0) return;
    ProfScope ps(ctx, "fz_merge");
    fz_merge_kernel<<<(unsigned)((dict_slots + 255) / 256), 256, 0, ctx.stream>>>(M, dict_slots);
    LAUNCH_CHECK(ctx);
}

}  // namespace auron
