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
//   fused   (one warp per tile)   predicate columns are unpacked to pass bits, the selection of the tile is built with
//                                 ballots; tiles without a selected row stop here.  Then the key column is unpacked to
//                                 accumulator slots and the argument columns to values, and the selected rows update the
//                                 accumulators with L2 reductions (RED).
//   merge   (per batch)           see below
//
// Aggregation in dictionary space: a dictionary-encoded group key is never looked up.  A row's slot is its dictionary
// INDEX (offset by the dictionary's base in a per-batch accumulator array), which removes the divergent 4-byte gather per
// row (the L1 wavefront limit, ~2 cycles per lane per SM: the 1.59 ms of round 1's ss_item_sk decode) from the row loop; after
// the batch one thread per dictionary entry folds its accumulators into the direct table at key - kmin.  PLAIN key pages
// address the direct table straight away.
//
// Roofline: HBM-bound on the encoded bytes (the only DRAM traffic that scales with rows); the row loop is bound by the SM's
// LSU issue rate for RED (1.29 cycles per lane per SM, B300_MICROARCH.md "Atomics").
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
    int seg = C.seg_base[page_id];
    int64_t v0 = 0;
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
constexpr int FZ_WARPS = 4;
struct FzSmemCol {
    uint32_t vals[FZ_TILE];   // per non-null value of the tile, in row order: pass bit / accumulator slot / argument value
    uint32_t w[32];           // validity words of the tile
    int32_t pref[32];         // non-null values before each word
};
struct FzSegCtx {
    int32_t role;
    bool dict;
    uint32_t ndict;
    const uint8_t* ddata;
    const uint32_t* pass;   // FZ_PRED: first pass word of this dictionary
    uint32_t slot_base;     // FZ_KEY
    int64_t lo, hi;
    long long kmin;
    int64_t range;
    int32_t* oor;
};
__device__ __forceinline__ uint32_t fz_xform(const FzSegCtx& x, uint32_t raw) {
    if (x.dict) {
        const uint32_t i = raw < x.ndict ? raw : 0u;   // corrupt index guard
        if (x.ndict == 0) return x.role == FZ_KEY ? x.slot_base : 0u;
        switch (x.role) {
            case FZ_PRED: return (x.pass[i >> 5] >> (i & 31)) & 1u;
            case FZ_KEY: return x.slot_base + i;
            default: return ld_u32_unaligned(x.ddata + (int64_t)i * 4);
        }
    }
    switch (x.role) {
        case FZ_PRED: {
            const int64_t v = (int64_t)(int32_t)raw;
            return (v >= x.lo && v <= x.hi) ? 1u : 0u;
        }
        case FZ_KEY: {
            int64_t s = (int64_t)(int32_t)raw - x.kmin;
            if ((uint64_t)s >= (uint64_t)x.range) {   // the column statistics did not cover this value
                *x.oor = 1;
                s = x.range;
            }
            return 0x80000000u | (uint32_t)s;
        }
        default: return raw;
    }
}
// the nvs non-null values of segment S, transformed for the column's role, to dst[0, nvs)
__device__ __forceinline__ void fz_unpack(const FzLaunch& L, const FzColumn& C, const FzSeg& S, const PqPage& pg, uint32_t* dst, int nvs, unsigned lane) {
    FzSegCtx x;
    x.role = C.role;
    x.dict = pg.encoding == 2 || pg.encoding == 8;
    x.lo = C.lo;
    x.hi = C.hi;
    x.kmin = L.kmin;
    x.range = L.range;
    x.oor = L.oor;
    x.ndict = 0;
    x.ddata = nullptr;
    x.pass = nullptr;
    x.slot_base = 0;
    const uint8_t* vals = pg.val_ptr;
    if (x.dict) {
        const PqDict dd = C.dicts[pg.dict_id];
        x.ndict = (uint32_t)dd.num_values;
        x.ddata = dd.data;
        if (C.role == FZ_PRED) x.pass = C.pass_bits + C.pass_off[pg.dict_id];
        if (C.role == FZ_KEY) x.slot_base = (uint32_t)C.dslot_base[pg.dict_id];
        const int bw = pg.val_len > 0 ? vals[0] : 0;
        Hybrid idx;
        hybrid_restore(idx, S.idx, vals + 1, vals + pg.val_len, bw);
        int pos = 0;
        while (pos < nvs) {
            if (idx.run_remaining == 0) idx.next_run();
            const int t = min(nvs - pos, idx.run_remaining);
            if (idx.is_rle) {
                const uint32_t o = fz_xform(x, idx.rle_value);
                for (int k = lane; k < t; k += 32) dst[pos + k] = o;
            } else {
                // lane L unpacks values L, L + 32, ...: 32 values are exactly `bw` 32-bit words, so the word pointer advances by
                // bw per step and the sub-word shift is a per-lane constant of the run
                const int64_t bit0 = (int64_t)(idx.bp_consumed + (int)lane) * bw;
                const uintptr_t qa = (uintptr_t)(idx.bp_base + (bit0 >> 3));
                const uint32_t* wp = (const uint32_t*)(qa & ~(uintptr_t)3);
                const unsigned sh = (unsigned)(qa & 3) * 8 + (unsigned)(bit0 & 7);
                const uint32_t vmask = bw >= 32 ? 0xffffffffu : ((1u << bw) - 1u);
                uint32_t* d = dst + pos + (int)lane;
                int k = lane;
                for (; k + 96 < t; k += 128) {   // 4 independent unpacks in flight per lane
                    uint32_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = __funnelshift_r(wp[u * bw], wp[u * bw + 1], sh) & vmask;
#pragma unroll
                    for (int u = 0; u < 4; u++) d[32 * u] = fz_xform(x, v[u]);
                    wp += 4 * bw;
                    d += 128;
                }
                for (; k < t; k += 32) {
                    *d = fz_xform(x, __funnelshift_r(wp[0], wp[1], sh) & vmask);
                    wp += bw;
                    d += 32;
                }
            }
            pos += t;
            idx.run_remaining -= t;
            if (!idx.is_rle) idx.bp_consumed += t;
        }
    } else {   // PLAIN INT32: value k of the segment is the 32-bit word at vals + 4 (v0 + k) (no alignment guarantee)
        const uintptr_t ba = (uintptr_t)(vals + S.v0 * 4);
        const uint32_t* bw32 = (const uint32_t*)(ba & ~(uintptr_t)3);
        const unsigned bsh = (unsigned)(ba & 3) * 8;
        for (int k = lane; k < nvs; k += 32) {
            const uint32_t raw = bsh ? __funnelshift_r(bw32[k], bw32[k + 1], bsh) : bw32[k];
            dst[k] = fz_xform(x, raw);
        }
    }
}
// validity words + rank bases + transformed values of role-column c for tile T
__device__ __forceinline__ void fz_load_col(const FzLaunch& L, int c, int T, int n_tile, FzSmemCol& sc, unsigned lane) {
    const FzColumn& C = L.col[c];
    const int cnt = n_tile - 32 * (int)lane;
    uint32_t w = cnt >= 32 ? 0xffffffffu : (cnt > 0 ? (1u << cnt) - 1u : 0u);
    if (C.valid) w &= C.valid[(int64_t)T * (FZ_TILE / 32) + lane];
    const int pc = __popc(w);
    int inc = pc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, inc, d);
        if ((int)lane >= d) inc += t;
    }
    sc.w[lane] = w;
    sc.pref[lane] = inc - pc;
    int seg = C.first_seg[T], covered = 0, pos = 0;
    for (int guard = 0; covered < n_tile && guard < FZ_TILE; guard++, seg++) {
        const FzSeg S = C.segs[seg];
        const PqPage pg = C.pages[S.page];
        const int nvs = min(S.nvalid, FZ_TILE - pos);
        if (nvs > 0) fz_unpack(L, C, S, pg, sc.vals + pos, nvs, lane);
        pos += max(nvs, 0);
        covered += S.n > 0 ? S.n : FZ_TILE;
    }
    __syncwarp();
}
__global__ void __launch_bounds__(FZ_WARPS * 32) fz_kernel(const __grid_constant__ FzLaunch L) {
    extern __shared__ __align__(16) uint8_t fz_smem[];
    const int wid = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    const int T = blockIdx.x * FZ_WARPS + wid;
    if (T >= L.n_tiles) return;
    FzSmemCol* sc = (FzSmemCol*)fz_smem + (size_t)wid * L.ncols;
    uint32_t* s_sel = (uint32_t*)((FzSmemCol*)fz_smem + (size_t)FZ_WARPS * L.ncols) + wid * 32;
    const int n_tile = (int)min((int64_t)FZ_TILE, L.n_rows - (int64_t)T * FZ_TILE);
    const unsigned lt = lanemask_lt();
    // ---- 1. predicate columns -> selection words of the tile (lane j ends up with the word of rows [32 j, 32 j + 32))
    for (int c = 0; c < L.npred; c++) fz_load_col(L, c, T, n_tile, sc[c], lane);
    uint32_t my_sel = 0;
    for (int j = 0; j * 32 < n_tile; j++) {
        bool pass = 32 * j + (int)lane < n_tile;
        for (int p = 0; p < L.npred; p++) {
            const uint32_t w = sc[p].w[j];
            const bool v = (w >> lane) & 1u;
            const uint32_t bit = v ? sc[p].vals[sc[p].pref[j] + __popc(w & lt)] : 0u;
            pass = pass && bit;
        }
        const uint32_t word = __ballot_sync(FULL_MASK, pass);
        if ((int)lane == j) my_sel = word;
    }
    int nsel = __popc(my_sel);
#pragma unroll
    for (int d = 16; d; d >>= 1) nsel += __shfl_xor_sync(FULL_MASK, nsel, d);
    if (nsel == 0) return;   // nothing of this tile survives the filter: the other columns are not even unpacked
    if (lane == 0) atomicAdd(L.selected_rows, (unsigned long long)nsel);
    s_sel[lane] = my_sel;
    // ---- 2. key and argument columns
    for (int c = L.npred; c < L.ncols; c++) fz_load_col(L, c, T, n_tile, sc[c], lane);
    __syncwarp();
    // ---- 3. selected rows -> accumulators
    const FzSmemCol& K = sc[L.key_col];
    for (int j = 0; j * 32 < n_tile; j++) {
        const uint32_t sw = s_sel[j];
        if (sw == 0) continue;
        if (!((sw >> lane) & 1u)) continue;
        const uint32_t wk = K.w[j];
        uint32_t U = 0x80000000u | (uint32_t)L.range;   // NULL key group
        if ((wk >> lane) & 1u) U = K.vals[K.pref[j] + __popc(wk & lt)];
        const bool dsp = !(U >> 31);
        const int64_t slot = (int64_t)(U & 0x7fffffffu);
        bool marked = false;
        for (int a = 0; a < L.nacc; a++) {
            const FzAcc& A = L.acc[a];
            bool ok = true;
            long long v = 1;
            if (A.col >= 0) {
                const uint32_t wv = sc[A.col].w[j];
                ok = (wv >> lane) & 1u;
                if (ok && A.kind != ACC_COUNT) v = (long long)(int32_t)sc[A.col].vals[sc[A.col].pref[j] + __popc(wv & lt)];
            }
            if (!ok) continue;
            unsigned long long* p = (dsp ? A.dspace : A.direct) + slot;
            switch (A.kind) {
                case ACC_MIN: atomicMin((long long*)p, v); break;
                case ACC_MAX: atomicMax((long long*)p, v); break;
                default: atomicAdd(p, (unsigned long long)v); break;   // SUM (wrapping, sum.rs:115) / COUNT
            }
            uint8_t* vb = dsp ? A.dspace_valid : A.direct_valid;
            if (vb) vb[slot] = 1;
            marked = marked || A.kind == ACC_COUNT || vb != nullptr;
        }
        if (!marked) (dsp ? L.seen_dspace : L.seen_direct)[slot] = 1;   // the group exists although no accumulator shows it
    }
}
static size_t fz_smem_bytes(int ncols) { return (size_t)FZ_WARPS * ncols * sizeof(FzSmemCol) + (size_t)FZ_WARPS * 32 * 4; }
void fz_run(Ctx& ctx, const FzLaunch& L) {
    if (L.n_tiles <= 0) return;
    static size_t attr_set = 0;
    const size_t smem = fz_smem_bytes(L.ncols);
    if (smem > attr_set) {
        CUDA_OK(cudaFuncSetAttribute(fz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fz_smem_bytes(FZ_MAX_COLS)));
        attr_set = fz_smem_bytes(FZ_MAX_COLS);
    }
    ProfScope ps(ctx, "fz_scan_filter_agg");
    fz_kernel<<<(L.n_tiles + FZ_WARPS - 1) / FZ_WARPS, FZ_WARPS * 32, smem, ctx.stream>>>(L);
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
    bool vb[FZ_MAX_ACCS];
#pragma unroll
    for (int a = 0; a < FZ_MAX_ACCS; a++) {
        v[a] = 0;
        vb[a] = false;
        if (a < M.nacc) {
            v[a] = M.acc[a].dspace[g];
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
