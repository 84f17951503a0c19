// k_hash.cu -- Spark-compatible murmur3 / xxhash64 over Arrow columns and the shuffle partition id
// (row S3 of SURVEY.md section 8a).  Replaces create_murmur3_hashes / create_xxhash64_hashes
// (datafusion-ext-commons/src/spark_hash.rs:28-224) and evaluate_partition_ids
// (datafusion-ext-plans/src/shuffle/mod.rs:163-188).
//
// HBM-bound: algorithmic bytes per row = sum of key value widths (+1/8 B validity) in, 4 B (ids) out.
// One fused kernel walks all key columns of a row in registers (the reference makes one pass per
// column over a hash buffer); fixed-width columns are read with coalesced per-thread loads in a
// grid-stride loop; utf8 rows are hashed one thread per row (neighbouring rows share cache lines).
#include "device_utils.cuh"
#include "kernels.h"

namespace auron {

struct HashCol {
    const void* data;
    const uint8_t* validity;
    const int32_t* offsets;
    int32_t type;
};
constexpr int kMaxHashCols = 16;
struct HashArgs {
    HashCol c[kMaxHashCols];
    int32_t ncols;
};

template <int KIND>
__device__ __forceinline__ uint64_t hash_one(const HashCol& col, int64_t row, uint64_t h) {
    if (col.validity && !bit_get(col.validity, row)) return h;  // NULL leaves the running hash (spark_hash.rs:78-84)
    switch (col.type) {
        case T_BOOL: {  // bool -> u32 1/0 (:131-158)
            uint32_t v = bit_get((const uint8_t*)col.data, row) ? 1u : 0u;
            return KIND == 0 ? (uint64_t)murmur3_u32(v, (uint32_t)h) : xxhash64_u32(v, h);
        }
        case T_INT8: {  // i8/i16 are widened to i32 (:160-165)
            uint32_t v = (uint32_t)(int32_t)((const int8_t*)col.data)[row];
            return KIND == 0 ? (uint64_t)murmur3_u32(v, (uint32_t)h) : xxhash64_u32(v, h);
        }
        case T_INT16: {
            uint32_t v = (uint32_t)(int32_t)((const int16_t*)col.data)[row];
            return KIND == 0 ? (uint64_t)murmur3_u32(v, (uint32_t)h) : xxhash64_u32(v, h);
        }
        case T_INT32: case T_DATE32: case T_FLOAT32: {
            uint32_t v = ((const uint32_t*)col.data)[row];
            return KIND == 0 ? (uint64_t)murmur3_u32(v, (uint32_t)h) : xxhash64_u32(v, h);
        }
        case T_INT64: case T_DATE64: case T_TIMESTAMP: case T_FLOAT64: {
            uint64_t v = ((const uint64_t*)col.data)[row];
            return KIND == 0 ? (uint64_t)murmur3_u64(v, (uint32_t)h) : xxhash64_u64(v, h);
        }
        case T_DECIMAL128: {  // 16 LE bytes (:110-129)
            ulonglong2 v = ((const ulonglong2*)col.data)[row];
            return KIND == 0 ? (uint64_t)murmur3_u128(v.x, v.y, (uint32_t)h) : xxhash64_u128(v.x, v.y, h);
        }
        case T_UTF8: case T_BINARY: {
            int32_t b = col.offsets[row], e = col.offsets[row + 1];
            const uint8_t* p = (const uint8_t*)col.data + b;
            return KIND == 0 ? (uint64_t)murmur3_bytes(p, e - b, (uint32_t)h) : xxhash64_bytes(p, e - b, h);
        }
        default: return h;
    }
}

// OUT_MODE 0: raw hash (int32 for murmur3, int64 for xxhash64); 1: pmod partition id (murmur3 only)
template <int KIND, int OUT_MODE>
__global__ void __launch_bounds__(256) hash_rows_kernel(HashArgs a, int64_t n, uint64_t seed, int32_t num_parts, void* __restrict__ out) {
    int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x; row < n; row += stride) {
        uint64_t h = seed;
        for (int c = 0; c < a.ncols; c++) h = hash_one<KIND>(a.c[c], row, h);
        if (KIND == 0) {
            int32_t hv = (int32_t)(uint32_t)h;
            if (OUT_MODE == 1) {
                int32_t r = hv % num_parts;   // rem_euclid (shuffle/mod.rs:178-188)
                if (r < 0) r += num_parts;
                hv = r;
            }
            ((int32_t*)out)[row] = hv;
        } else {
            ((int64_t*)out)[row] = (int64_t)h;
        }
    }
}

// Fixed-width keys (the shuffle / join keys of TPC-DS are int32 / int64 surrogate keys): four consecutive rows per thread, one
// 128-bit load per 4-byte column (two per 8-byte column), one validity nibble, one 128-bit store of the four results; the
// partition id is Lemire's fastmod (two multiplies, exact for 32-bit operands) instead of a runtime integer division.
// Round 1's row-per-thread kernel reached 27.8 % of the HBM peak on 64M int32 keys.
struct FastMod {
    uint64_t m;        // floor(2^64 / d) + 1
    uint32_t d, c;     // divisor, 2^31 mod d
};
__device__ __forceinline__ uint32_t fastmod_u32(uint32_t v, const FastMod& f) { return (uint32_t)__umul64hi(f.m * (uint64_t)v, (uint64_t)f.d); }
// ((int32)hv).rem_euclid(d) (shuffle/mod.rs:178-188) without a signed division: hv + 2^31 is non-negative
__device__ __forceinline__ int32_t pmod_i32(int32_t hv, const FastMod& f) {
    const uint32_t r = fastmod_u32((uint32_t)hv ^ 0x80000000u, f);
    return (int32_t)(r >= f.c ? r - f.c : r + f.d - f.c);
}
template <int KIND, int OUT_MODE>
__global__ void __launch_bounds__(256) hash_fixed4_kernel(HashArgs a, int64_t n, uint64_t seed, FastMod fm, void* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    for (int64_t r0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; r0 < n; r0 += stride) {
        uint64_t h[4] = {seed, seed, seed, seed};
        const bool full = r0 + 4 <= n;
        for (int c = 0; c < a.ncols; c++) {
            const HashCol& col = a.c[c];
            const uint32_t vm = col.validity ? (uint32_t)(col.validity[r0 >> 3] >> (r0 & 4)) & 0xFu : 0xFu;   // r0 is a multiple of 4
            const bool w8 = col.type == T_INT64 || col.type == T_DATE64 || col.type == T_TIMESTAMP || col.type == T_FLOAT64;
            if (!w8) {
                uint32_t v[4];
                if (full) {
                    const uint4 q = *(const uint4*)((const uint32_t*)col.data + r0);
                    v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) v[k] = r0 + k < n ? ((const uint32_t*)col.data)[r0 + k] : 0u;
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((vm >> k) & 1u) h[k] = KIND == 0 ? (uint64_t)murmur3_u32(v[k], (uint32_t)h[k]) : xxhash64_u32(v[k], h[k]);
            } else {
                uint64_t v[4];
                if (full) {
                    const ulonglong2 q0 = *(const ulonglong2*)((const uint64_t*)col.data + r0), q1 = *(const ulonglong2*)((const uint64_t*)col.data + r0 + 2);
                    v[0] = q0.x, v[1] = q0.y, v[2] = q1.x, v[3] = q1.y;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) v[k] = r0 + k < n ? ((const uint64_t*)col.data)[r0 + k] : 0ull;
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if ((vm >> k) & 1u) h[k] = KIND == 0 ? (uint64_t)murmur3_u64(v[k], (uint32_t)h[k]) : xxhash64_u64(v[k], h[k]);
            }
        }
        if (KIND == 0) {
            int32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = OUT_MODE == 1 ? pmod_i32((int32_t)(uint32_t)h[k], fm) : (int32_t)(uint32_t)h[k];
            if (full) *(int4*)((int32_t*)out + r0) = make_int4(o[0], o[1], o[2], o[3]);
            else
                for (int k = 0; k < 4 && r0 + k < n; k++) ((int32_t*)out)[r0 + k] = o[k];
        } else {
            for (int k = 0; k < 4 && r0 + k < n; k++) ((int64_t*)out)[r0 + k] = (int64_t)h[k];
        }
    }
}
static bool all_fixed_4_or_8(const std::vector<ColumnPtr>& cols) {
    if (cols.empty() || getenv("AURON_DISABLE_HASH_FIXED4")) return false;
    for (auto& c : cols) {
        switch (c->type.id) {
            case T_INT32: case T_DATE32: case T_FLOAT32: case T_INT64: case T_DATE64: case T_TIMESTAMP: case T_FLOAT64: break;
            default: return false;
        }
        if (!c->data || ((uintptr_t)c->data->ptr & 15)) return false;   // 128-bit loads
    }
    return true;
}
static FastMod make_fastmod(int32_t d) {
    FastMod f;
    f.d = (uint32_t)d;
    f.m = ~0ull / (uint64_t)f.d + 1;
    f.c = (uint32_t)((1ull << 31) % f.d);
    return f;
}

static HashArgs make_args(const std::vector<ColumnPtr>& cols) {
    AURON_CHECK((int)cols.size() <= kMaxHashCols, "too many hash columns");
    HashArgs a;
    a.ncols = (int)cols.size();
    for (int i = 0; i < a.ncols; i++) {
        a.c[i].data = cols[i]->data ? cols[i]->data->ptr : nullptr;
        a.c[i].validity = cols[i]->vbits();
        a.c[i].offsets = P<int32_t>(cols[i]->offsets);
        a.c[i].type = cols[i]->type.id;
        if (cols[i]->type.id == T_NULL) a.c[i].type = T_NULL;
    }
    return a;
}

static unsigned grid_for(Ctx& ctx, int64_t n) {
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)ctx.sm_count * 16;   // 16 x 256-thread CTAs per SM = 2 full waves of resident warps
    return (unsigned)std::max<int64_t>(1, std::min(blocks, cap));
}

Buf hash_columns(Ctx& ctx, const std::vector<ColumnPtr>& cols, int64_t n, int kind, int64_t seed) {
    HashArgs a = make_args(cols);
    Buf out = dalloc(ctx, (size_t)std::max<int64_t>(n, 1) * (kind == 0 ? 4 : 8));
    if (n == 0) return out;
    const bool fixed = all_fixed_4_or_8(cols);
    if (kind == 0 && fixed) hash_fixed4_kernel<0, 0><<<grid_for(ctx, (n + 3) / 4), 256, 0, ctx.stream>>>(a, n, (uint64_t)(uint32_t)(int32_t)seed, make_fastmod(1), out->ptr);
    else if (kind == 0) hash_rows_kernel<0, 0><<<grid_for(ctx, n), 256, 0, ctx.stream>>>(a, n, (uint64_t)(uint32_t)(int32_t)seed, 1, out->ptr);
    else if (fixed) hash_fixed4_kernel<1, 0><<<grid_for(ctx, (n + 3) / 4), 256, 0, ctx.stream>>>(a, n, (uint64_t)seed, make_fastmod(1), out->ptr);
    else hash_rows_kernel<1, 0><<<grid_for(ctx, n), 256, 0, ctx.stream>>>(a, n, (uint64_t)seed, 1, out->ptr);
    CUDA_OK(cudaGetLastError());
    launch_count(ctx);
    return out;
}

// RoundRobinPartitioning: row i of the chunk goes to (i + start) % num_parts (evaluate_robin_partition_ids, shuffle/mod.rs:190-202)
__global__ void round_robin_ids_kernel(int32_t* __restrict__ out, int64_t n, int64_t start, int32_t num_parts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)((i + start) % num_parts);
}
Buf round_robin_partition_ids(Ctx& ctx, int64_t n, int64_t start, int32_t num_parts) {
    Buf out = dalloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4);
    if (n > 0) {
        round_robin_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx.stream>>>(P<int32_t>(out), n, start, num_parts);
        CUDA_OK(cudaGetLastError());
        launch_count(ctx);
    }
    return out;
}
Buf murmur3_partition_ids(Ctx& ctx, const std::vector<ColumnPtr>& cols, int64_t n, int32_t num_parts, int32_t seed) {
    AURON_CHECK(num_parts > 0, "num_parts must be positive");
    HashArgs a = make_args(cols);
    Buf out = dalloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4);
    if (n == 0) return out;
    ProfScope ps(ctx, "murmur3_partition_ids");
    if (all_fixed_4_or_8(cols)) hash_fixed4_kernel<0, 1><<<grid_for(ctx, (n + 3) / 4), 256, 0, ctx.stream>>>(a, n, (uint64_t)(uint32_t)seed, make_fastmod(num_parts), out->ptr);
    else hash_rows_kernel<0, 1><<<grid_for(ctx, n), 256, 0, ctx.stream>>>(a, n, (uint64_t)(uint32_t)seed, num_parts, out->ptr);
    CUDA_OK(cudaGetLastError());
    launch_count(ctx);
    return out;
}

}  // namespace auron
