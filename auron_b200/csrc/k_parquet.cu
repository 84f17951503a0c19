// k_parquet.cu -- Parquet page decode on device (row P1 of SURVEY.md section 8a): RLE / bit-packed hybrid
// definition levels and dictionary indices, dictionary gather, PLAIN fixed-width values, PLAIN byte arrays,
// NULL scatter from definition levels and the INT32/INT64 -> Decimal128 widening of AuronSchemaAdapter
// (datafusion-ext-plans/src/scan/mod.rs:103-160).  The reference does this on the CPU inside the third-party
// `parquet` crate (call site parquet_exec.rs:175-197).
//
// One warp decodes one data page in a single pass: the hybrid streams are parsed run by run (the run header
// is read redundantly by all lanes = one broadcast load), each 32-row chunk turns its definition levels into
// a validity word with a ballot, the ranks of the valid lanes (popc of the lower-lane mask) index the value
// stream, values are fetched from the dictionary (L2-resident) or the PLAIN section and written once,
// converted to the Arrow type.  Pages of all row groups of a batch are decoded by one launch per column.
// HBM-bound: algorithmic bytes = encoded page bytes in + Arrow bytes out (+1/8 B validity).
#include "device_utils.cuh"
#include "kernels.h"
#include "parquet_dev.h"

namespace auron {

#define LAUNCH_CHECK(ctx)            \
    do {                             \
        CUDA_OK(cudaGetLastError()); \
        launch_count(ctx);           \
    } while (0)

__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* p) {
    uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
__device__ __forceinline__ uint64_t ld_u64_unaligned(const uint8_t* p) {
    return (uint64_t)ld_u32_unaligned(p) | ((uint64_t)ld_u32_unaligned(p + 4) << 32);
}

// RLE / bit-packed hybrid stream, state replicated in every lane of the warp
struct Hybrid {
    const uint8_t* p;
    const uint8_t* end;
    int bw;
    int run_remaining;
    bool is_rle;
    uint32_t rle_value;
    const uint8_t* bp_base;
    int bp_consumed;

    __device__ void init(const uint8_t* b, const uint8_t* e, int bit_width) {
        p = b;
        end = e;
        bw = bit_width;
        run_remaining = 0;
        is_rle = true;
        rle_value = 0;
        bp_base = b;
        bp_consumed = 0;
    }
    __device__ void next_run() {
        uint32_t h = 0;
        int shift = 0;
        while (p < end) {
            uint8_t b = *p++;
            h |= (uint32_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) break;
            shift += 7;
        }
        if (h & 1) {
            int groups = (int)(h >> 1);
            is_rle = false;
            run_remaining = groups * 8;
            bp_base = p;
            bp_consumed = 0;
            p += (int64_t)groups * bw;
        } else {
            is_rle = true;
            run_remaining = (int)(h >> 1);
            int nb = (bw + 7) / 8;
            uint32_t v = 0;
            for (int i = 0; i < nb && p + i < end; i++) v |= (uint32_t)p[i] << (8 * i);
            rle_value = v;
            p += nb;
        }
        if (run_remaining == 0 && p >= end) run_remaining = 1 << 30;   // exhausted stream: pad with the last value (malformed input guard)
    }
    // lane j < m receives the j-th next value of the stream
    __device__ uint32_t read_batch(int m, unsigned lane) {
        uint32_t v = 0;
        int filled = 0;
        while (filled < m) {
            if (run_remaining == 0) next_run();
            int t = min(m - filled, run_remaining);
            if ((int)lane >= filled && (int)lane < filled + t) {
                if (is_rle) v = rle_value;
                else {
                    int64_t bitpos = (int64_t)(bp_consumed + ((int)lane - filled)) * bw;
                    const uint8_t* q = bp_base + (bitpos >> 3);
                    uint64_t w = 0;
                    int nb = (int)((bitpos & 7) + bw + 7) / 8;
                    for (int i = 0; i < nb; i++) w |= (uint64_t)q[i] << (8 * i);
                    v = (uint32_t)((w >> (bitpos & 7)) & ((bw == 32) ? 0xffffffffull : ((1ull << bw) - 1ull)));
                }
            }
            filled += t;
            run_remaining -= t;
            if (!is_rle) bp_consumed += t;
        }
        return v;
    }
};

__device__ __forceinline__ void store_converted(const PqColumnArgs& a, const uint8_t* src, int64_t row) {
    // src points at one physical value (little-endian INT32/INT64/FLOAT/DOUBLE, big-endian FLBA)
    switch (a.phys_type) {
        case 1: case 4: {   // INT32 / FLOAT
            uint32_t v = ld_u32_unaligned(src);
            switch (a.out_type) {
                case T_INT8: ((int8_t*)a.out)[row] = (int8_t)v; break;
                case T_INT16: ((int16_t*)a.out)[row] = (int16_t)v; break;
                case T_INT32: case T_DATE32: case T_FLOAT32: ((uint32_t*)a.out)[row] = v; break;
                case T_INT64: case T_TIMESTAMP: case T_DATE64: ((int64_t*)a.out)[row] = (int64_t)(int32_t)v; break;
                case T_FLOAT64: ((double*)a.out)[row] = (double)__int_as_float((int)v); break;
                case T_DECIMAL128: {   // scan/mod.rs:131-136: value copy, no rescale
                    int64_t s = (int64_t)(int32_t)v;
                    ((int64_t*)a.out)[2 * row] = s;
                    ((int64_t*)a.out)[2 * row + 1] = s < 0 ? -1 : 0;
                    break;
                }
            }
            break;
        }
        case 2: case 5: {   // INT64 / DOUBLE
            uint64_t v = ld_u64_unaligned(src);
            switch (a.out_type) {
                case T_INT32: case T_DATE32: ((int32_t*)a.out)[row] = (int32_t)v; break;
                case T_DECIMAL128:
                    ((uint64_t*)a.out)[2 * row] = v;
                    ((int64_t*)a.out)[2 * row + 1] = ((int64_t)v) < 0 ? -1 : 0;
                    break;
                default: ((uint64_t*)a.out)[row] = v; break;
            }
            break;
        }
        case 7: {   // FIXED_LEN_BYTE_ARRAY decimal: big-endian two's complement
            int n = a.type_length;
            uint64_t hi = (src[0] & 0x80) ? ~0ull : 0ull, lo = hi;
            for (int i = 0; i < n; i++) {
                hi = (hi << 8) | (lo >> 56);
                lo = (lo << 8) | src[i];
            }
            if (a.out_type == T_DECIMAL128) {
                ((uint64_t*)a.out)[2 * row] = lo;
                ((uint64_t*)a.out)[2 * row + 1] = hi;
            } else if (a.out_type == T_INT64) ((uint64_t*)a.out)[row] = lo;
            else ((uint32_t*)a.out)[row] = (uint32_t)lo;
            break;
        }
    }
}
__device__ __forceinline__ void store_zero(const PqColumnArgs& a, int64_t row) {
    switch (a.out_width) {
        case 1: ((uint8_t*)a.out)[row] = 0; break;
        case 2: ((uint16_t*)a.out)[row] = 0; break;
        case 4: ((uint32_t*)a.out)[row] = 0; break;
        case 8: ((uint64_t*)a.out)[row] = 0; break;
        case 16: ((uint64_t*)a.out)[2 * row] = 0; ((uint64_t*)a.out)[2 * row + 1] = 0; break;
    }
}

constexpr int PQ_WARPS = 4;

__global__ void __launch_bounds__(PQ_WARPS * 32) pq_decode_pages_kernel(PqColumnArgs a) {
    int page_id = blockIdx.x * PQ_WARPS + (threadIdx.x >> 5);
    if (page_id >= a.n_pages) return;
    const unsigned lane = lane_id();
    const PqPage pg = a.pages[page_id];
    Hybrid def, idx;
    const bool has_def = a.max_def > 0 && pg.def_len > 0;
    if (has_def) def.init(pg.def_ptr, pg.def_ptr + pg.def_len, 1);
    const bool dict = pg.encoding == 2 || pg.encoding == 8;
    const uint8_t* vals = pg.val_ptr;
    const bool bool_rle = a.phys_type == 0 && pg.encoding == 3;   // RLE booleans (data page v2 writers): u32 length + hybrid, bit width 1
    if (dict) {
        int bw = pg.val_len > 0 ? vals[0] : 0;
        idx.init(vals + 1, vals + pg.val_len, bw);
    } else if (bool_rle) {
        idx.init(vals + 4, vals + pg.val_len, 1);
    }
    const PqDict dd = dict ? a.dicts[pg.dict_id] : PqDict{nullptr, 0, 0};
    const int w = a.phys_width;
    int done = 0;
    int64_t value_base = 0;
    const int rows = pg.num_values;
    while (done < rows) {
        int64_t out_row = (int64_t)pg.row_start + done;
        int m = min(rows - done, 32 - (int)(out_row & 31));
        bool active = (int)lane < m;
        bool valid = active;
        if (has_def) {
            uint32_t d = def.read_batch(m, lane);
            valid = active && d == (uint32_t)a.max_def;
        } else if (a.max_def > 0 && pg.all_null) valid = false;
        uint32_t mask = __ballot_sync(FULL_MASK, valid);
        int rank = __popc(mask & lanemask_lt());
        int nvalid = __popc(mask);
        int64_t ordinal = value_base + rank;
        uint32_t myidx = 0;
        if (dict) {
            uint32_t got = nvalid ? idx.read_batch(nvalid, lane) : 0;
            myidx = __shfl_sync(FULL_MASK, got, rank & 31);
            if (myidx >= (uint32_t)dd.num_values) myidx = 0;   // corrupt index guard
        }
        int64_t row = out_row + lane;
        if (active) {
            if (a.mode == PQ_MODE_INDEX) {
                // strings: position in the chunk's value table (dictionary entries first, then PLAIN values)
                a.out_idx[row] = valid ? (dict ? dd.value_base + (int32_t)myidx : pg.plain_value_base + (int32_t)ordinal) : -1;
            } else if (a.phys_type == 0) {
                // BOOLEAN PLAIN = LSB-first bit-packed without run headers; handled below through a ballot
            } else if (valid) {
                const uint8_t* src = dict ? dd.data + (int64_t)myidx * w : vals + ordinal * w;
                store_converted(a, src, row);
            } else {
                store_zero(a, row);
            }
        }
        if (a.phys_type == 0 && a.mode != PQ_MODE_INDEX) {
            bool bit = false;
            if (bool_rle) {
                uint32_t got = nvalid ? idx.read_batch(nvalid, lane) : 0;
                uint32_t mine = __shfl_sync(FULL_MASK, got, rank & 31);   // every lane takes part (no short-circuit around the shuffle)
                bit = valid && (mine & 1);
            } else if (valid) bit = (vals[ordinal >> 3] >> (ordinal & 7)) & 1;
            uint32_t bits = __ballot_sync(FULL_MASK, bit) << (out_row & 31);
            if (lane == 0 && bits) atomicOr(&((uint32_t*)a.out)[out_row >> 5], bits);
        }
        if (a.out_valid && lane == 0) {
            uint32_t bits = mask << (out_row & 31);
            if (m == 32) a.out_valid[out_row >> 5] = bits;
            else if (bits) atomicOr(&a.out_valid[out_row >> 5], bits);
        }
        done += m;
        value_base += nvalid;
    }
}

void pq_decode_pages(Ctx& ctx, const PqColumnArgs& a) {
    if (a.n_pages == 0) return;
    pq_decode_pages_kernel<<<(a.n_pages + PQ_WARPS - 1) / PQ_WARPS, PQ_WARPS * 32, 0, ctx.stream>>>(a);
    LAUNCH_CHECK(ctx);
}

// ---- PLAIN BYTE_ARRAY sections (dictionary pages and non-dictionary data pages): one thread walks one section
__global__ void pq_walk_byte_arrays_kernel(const PqByteSection* __restrict__ secs, int n_secs, int64_t* __restrict__ lens,
                                           const uint8_t** __restrict__ srcs) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_secs) return;
    PqByteSection sec = secs[s];
    const uint8_t* p = sec.ptr;
    const uint8_t* end = sec.ptr + sec.len;
    for (int i = 0; i < sec.num_values; i++) {
        uint32_t l = 0;
        if (p + 4 <= end) {
            l = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
            p += 4;
            if (p + l > end) l = (uint32_t)(end - p);
        }
        lens[sec.value_base + i] = l;
        srcs[sec.value_base + i] = p;
        p += l;
    }
}
__global__ void pq_copy_byte_arrays_kernel(const uint8_t* const* __restrict__ srcs, const int64_t* __restrict__ offs, int64_t n,
                                           int32_t* __restrict__ out_off, uint8_t* __restrict__ out_data) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    out_off[i] = (int32_t)offs[i];
    if (i == n) return;
    const uint8_t* s = srcs[i];
    uint8_t* d = out_data + offs[i];
    int64_t l = offs[i + 1] - offs[i];
    for (int64_t k = 0; k < l; k++) d[k] = s[k];
}

ColumnPtr pq_build_value_table(Ctx& ctx, const std::vector<PqByteSection>& secs, int64_t total_values, const DType& type) {
    auto col = std::make_shared<Column>();
    col->type = type;
    col->len = total_values;
    col->offsets = dalloc_zero(ctx, (size_t)(total_values + 1) * 4);
    if (total_values == 0 || secs.empty()) {
        col->data = dalloc(ctx, 0);
        return col;
    }
    Buf dsecs = to_device(ctx, secs.data(), secs.size() * sizeof(PqByteSection));
    Buf lens = dalloc_zero(ctx, (size_t)(total_values + 1) * 8);
    Buf srcs = dalloc_zero(ctx, (size_t)(total_values + 1) * 8);
    int n = (int)secs.size();
    pq_walk_byte_arrays_kernel<<<(n + 63) / 64, 64, 0, ctx.stream>>>(P<PqByteSection>(dsecs), n, P<int64_t>(lens), (const uint8_t**)srcs->ptr);
    LAUNCH_CHECK(ctx);
    exclusive_scan_i64(ctx, P<int64_t>(lens), P<int64_t>(lens), total_values, P<int64_t>(lens) + total_values);
    int64_t total = 0;
    to_host(ctx, &total, P<int64_t>(lens) + total_values, 8);
    AURON_CHECK(total <= (int64_t)INT32_MAX, "parquet string chunk exceeds 2 GiB");
    col->data = dalloc(ctx, (size_t)total);
    col->data_bytes = total;
    pq_copy_byte_arrays_kernel<<<(unsigned)((total_values + 1 + 255) / 256), 256, 0, ctx.stream>>>((const uint8_t* const*)srcs->ptr, P<int64_t>(lens), total_values,
                                                                                                 P<int32_t>(col->offsets), P<uint8_t>(col->data));
    LAUNCH_CHECK(ctx);
    ctx.sync();   // secs (host vector) was read by an async copy
    return col;
}

}  // namespace auron
