// k_serde.cu -- Auron's compacted shuffle / spill batch format on device (row S5 of SURVEY.md section 8a).
// Byte layout restated from datafusion-ext-commons/src/io/batch_serde.rs:
//   batch  := varint num_rows | column*                                        (:68-79)
//   column := varint has_nulls(0|1) | [validity bits re-packed from bit 0] | values
//   fixed width w > 1: byte-plane transposed values (all byte-0s, then byte-1s, ...)   (:273-307, transpose :292-305)
//   bool: value bits                                                            (:557-577)
//   utf8 / binary: lengths as i32 (transposed 4 x n) | concatenated bytes       (:603-633, :219-242)
// The input batch is already partition-contiguous (rows of partition p = [row_off[p], row_off[p+1])), so one launch
// per column serialises every partition at once: a thread owns a row, finds its partition by binary search over the
// row offsets, and scatters its bytes into the partition's planes (each plane write is coalesced across the warp).
// HBM-bound: algorithmic bytes = batch bytes in + serialized bytes out.
#include "device_utils.cuh"
#include "kernels.h"

namespace auron {

#define LAUNCH_CHECK(ctx)            \
    do {                             \
        CUDA_OK(cudaGetLastError()); \
        launch_count(ctx);           \
    } while (0)

struct SerSeg {
    int64_t out_validity;   // -1: column has no validity section
    int64_t out_values;     // fixed: planes ; bool: bits ; utf8: length planes
    int64_t out_bytes;      // utf8: payload
    int64_t byte_begin;     // utf8: source byte offset of the partition's first row
    int64_t row_begin;
    int64_t n;
};

__device__ __forceinline__ int find_part(const int64_t* __restrict__ row_off, int num_parts, int64_t row) {
    int lo = 0, hi = num_parts;   // largest p with row_off[p] <= row
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (row_off[mid] <= row) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) serde_column_kernel(const uint8_t* __restrict__ data, const uint8_t* __restrict__ validity,
                                                           const int32_t* __restrict__ offsets, int width, int is_bool, int is_varlen,
                                                           const int64_t* __restrict__ row_off, int num_parts, const SerSeg* __restrict__ segs,
                                                           int64_t n_rows, uint8_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rows) return;
    int p = find_part(row_off, num_parts, i);
    SerSeg s = segs[p];
    int64_t j = i - s.row_begin;
    if ((j & 7) == 0) {   // this thread emits the validity / bool byte of rows [i, i+8)
        int lim = (int)min((int64_t)8, s.n - j);
        if (s.out_validity >= 0) {
            uint8_t b = 0;
            for (int k = 0; k < lim; k++) b |= (uint8_t)(bit_get(validity, i + k) ? (1u << k) : 0u);
            out[s.out_validity + (j >> 3)] = b;
        }
        if (is_bool) {
            uint8_t b = 0;
            for (int k = 0; k < lim; k++) b |= (uint8_t)(bit_get(data, i + k) ? (1u << k) : 0u);
            out[s.out_values + (j >> 3)] = b;
        }
    }
    if (is_bool) return;
    if (is_varlen) {
        uint32_t len = (uint32_t)(offsets[i + 1] - offsets[i]);
#pragma unroll
        for (int k = 0; k < 4; k++) out[s.out_values + (int64_t)k * s.n + j] = (uint8_t)(len >> (8 * k));
        return;
    }
    const uint8_t* src = data + i * width;
    if (width == 1) out[s.out_values + j] = src[0];
    else
        for (int k = 0; k < width; k++) out[s.out_values + (int64_t)k * s.n + j] = src[k];
}
__global__ void __launch_bounds__(256) serde_bytes_kernel(const uint8_t* __restrict__ data, const int64_t* __restrict__ byte_off /*[num_parts+1]*/,
                                                          int num_parts, const SerSeg* __restrict__ segs, int64_t n_bytes, uint8_t* __restrict__ out) {
    int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (b >= n_bytes) return;
    int p = find_part(byte_off, num_parts, b);
    SerSeg s = segs[p];
    out[s.out_bytes + (b - s.byte_begin)] = data[b];
}
__global__ void gather_offsets_kernel(const int32_t* __restrict__ offsets, const int64_t* __restrict__ row_off, int n, int64_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = offsets[row_off[i]];
}

static int varint_len(uint64_t v) {
    int n = 1;
    while (v >= 128) {
        v /= 128;
        n++;
    }
    return n;
}
static int put_varint(uint64_t v, uint8_t* out) {   // io/mod.rs:61-69
    int n = 0;
    while (v >= 128) {
        out[n++] = (uint8_t)(128 + v % 128);
        v /= 128;
    }
    out[n++] = (uint8_t)v;
    return n;
}

SerializedParts serialize_partitions(Ctx& ctx, const Batch& b, const std::vector<int64_t>& row_offsets) {
    ProfScope ps_fn(ctx, "serde_write");
    const int num_parts = (int)row_offsets.size() - 1;
    const int ncols = (int)b.cols.size();
    SerializedParts res;
    res.part_offsets.assign(num_parts + 1, 0);
    Buf d_row_off = to_device(ctx, row_offsets.data(), row_offsets.size() * 8);
    // utf8 columns: byte offsets at partition boundaries
    std::vector<std::vector<int64_t>> byte_off(ncols);
    for (int c = 0; c < ncols; c++) {
        if (!b.cols[c]->type.is_varlen()) continue;
        Buf tmp = dalloc(ctx, (size_t)(num_parts + 1) * 8);
        gather_offsets_kernel<<<(num_parts + 1 + 255) / 256, 256, 0, ctx.stream>>>(P<int32_t>(b.cols[c]->offsets), P<int64_t>(d_row_off), num_parts + 1, P<int64_t>(tmp));
        LAUNCH_CHECK(ctx);
        byte_off[c].resize(num_parts + 1);
        to_host(ctx, byte_off[c].data(), tmp->ptr, (size_t)(num_parts + 1) * 8);
    }
    // layout
    std::vector<std::vector<SerSeg>> segs(ncols, std::vector<SerSeg>(num_parts));
    std::vector<std::pair<int64_t, std::vector<uint8_t>>> small;   // (offset, bytes) written from the host: varints
    int64_t pos = 0;
    for (int p = 0; p < num_parts; p++) {
        res.part_offsets[p] = pos;
        int64_t n = row_offsets[p + 1] - row_offsets[p];
        for (int c = 0; c < ncols; c++) segs[c][p] = SerSeg{-1, 0, 0, 0, row_offsets[p], n};
        if (n == 0) continue;   // empty partitions write nothing (ipc_compression.rs:68-70)
        std::vector<uint8_t> hdr(10);
        hdr.resize(put_varint((uint64_t)n, hdr.data()));
        small.emplace_back(pos, hdr);
        pos += varint_len((uint64_t)n);
        for (int c = 0; c < ncols; c++) {
            const Column& col = *b.cols[c];
            SerSeg& s = segs[c][p];
            if (col.type.id == T_NULL) continue;
            bool has_nulls = col.may_have_nulls();
            small.emplace_back(pos, std::vector<uint8_t>{(uint8_t)(has_nulls ? 1 : 0)});
            pos += 1;
            if (has_nulls) {
                s.out_validity = pos;
                pos += (n + 7) / 8;
            }
            s.out_values = pos;
            if (col.type.id == T_BOOL) pos += (n + 7) / 8;
            else if (col.type.is_varlen()) {
                pos += 4 * n;
                s.out_bytes = pos;
                s.byte_begin = byte_off[c][p];
                pos += byte_off[c][p + 1] - byte_off[c][p];
            } else pos += (int64_t)col.type.width() * n;
        }
    }
    res.part_offsets[num_parts] = pos;
    res.bytes = dalloc(ctx, (size_t)pos);
    uint8_t* out = P<uint8_t>(res.bytes);
    for (auto& kv : small) CUDA_OK(cudaMemcpyAsync(out + kv.first, kv.second.data(), kv.second.size(), cudaMemcpyHostToDevice, ctx.stream));
    int64_t n_rows = b.num_rows;
    for (int c = 0; c < ncols; c++) {
        const Column& col = *b.cols[c];
        if (col.type.id == T_NULL || n_rows == 0) continue;
        Buf dsegs = to_device(ctx, segs[c].data(), segs[c].size() * sizeof(SerSeg));
        serde_column_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, ctx.stream>>>(P<uint8_t>(col.data), col.vbits(), P<int32_t>(col.offsets), col.type.width(),
                                                                                   col.type.id == T_BOOL, col.type.is_varlen(), P<int64_t>(d_row_off), num_parts,
                                                                                   P<SerSeg>(dsegs), n_rows, out);
        LAUNCH_CHECK(ctx);
        if (col.type.is_varlen() && col.data_bytes > 0) {
            Buf dbo = to_device(ctx, byte_off[c].data(), byte_off[c].size() * 8);
            serde_bytes_kernel<<<(unsigned)((col.data_bytes + 255) / 256), 256, 0, ctx.stream>>>(P<uint8_t>(col.data), P<int64_t>(dbo), num_parts, P<SerSeg>(dsegs),
                                                                                              col.data_bytes, out);
            LAUNCH_CHECK(ctx);
        }
        ctx.sync();   // host vectors above are read by async copies
    }
    ctx.sync();
    return res;
}

BatchPtr deserialize_batch(Ctx&, const Schema&, const uint8_t*, int64_t, int64_t*) { fail("deserialize_batch: shuffle read is a 'next' row (SURVEY.md section 8f)"); }

}  // namespace auron
