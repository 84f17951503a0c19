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

// ---- fixed-width columns (2 / 4 / 8 / 16 bytes): one warp per 128-row chunk OF A PARTITION (chunks never straddle
// partitions and start on a byte boundary of the partition's validity section).  A lane takes 4 consecutive rows, packs
// their k-th bytes into one word per plane (byte_perm), and the warp writes each plane's 128 bytes as aligned 32-bit words:
// the plane start is not 4-byte aligned in general (varint headers, row counts), so lane q stores the word that straddles
// lanes q-1 and q (one shuffle + one funnel shift); the first and last few bytes of the run are stored byte-wise.
// The per-row kernel above wrote one byte per lane per store (32 B per warp instruction, 7.6 ms for 64M x 28 B).
__device__ __forceinline__ uint32_t pack_plane(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, int k) {
    const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);
    return __byte_perm(__byte_perm(x0, x1, sel), __byte_perm(x2, x3, sel), 0x5410);
}
template <int W>
__global__ void __launch_bounds__(256) serde_fixed_kernel(const uint8_t* __restrict__ data, const uint8_t* __restrict__ validity,
                                                          const int32_t* __restrict__ chunk_base, int num_parts, const SerSeg* __restrict__ segs,
                                                          int total_chunks, uint8_t* __restrict__ out) {
    const int chunk = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5);
    if (chunk >= total_chunks) return;
    const unsigned lane = threadIdx.x & 31;
    int lo = 0, hi = num_parts;   // largest p with chunk_base[p] <= chunk
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (chunk_base[mid] <= chunk) lo = mid;
        else hi = mid;
    }
    const SerSeg s = segs[lo];
    const int64_t j0 = (int64_t)(chunk - chunk_base[lo]) * 128;
    const int m = (int)min((int64_t)128, s.n - j0);
    const int64_t i0 = s.row_begin + j0 + 4 * lane;   // this lane's first source row
    const int mine = max(0, min(4, m - 4 * (int)lane));
    if (s.out_validity >= 0) {
        uint32_t nib = 0;
        if (mine > 0) {
            const uint32_t two = (uint32_t)validity[i0 >> 3] | ((uint32_t)validity[(i0 >> 3) + 1] << 8);   // bitmaps are padded (64 B)
            nib = (two >> (i0 & 7)) & ((1u << mine) - 1u);
        }
        const uint32_t up = __shfl_down_sync(FULL_MASK, nib, 1);
        if (!(lane & 1) && mine > 0) out[s.out_validity + (j0 >> 3) + (lane >> 1)] = (uint8_t)(nib | (up << 4));
    }
    constexpr int NW = W >= 4 ? W / 4 : 1;   // 32-bit words per row
    uint32_t x[4][NW];
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int q = 0; q < NW; q++) x[r][q] = 0;
        if (r < mine) {
            if (W == 2) x[r][0] = ((const uint16_t*)data)[i0 + r];
            else {
                const uint32_t* src = (const uint32_t*)data + (i0 + r) * NW;
#pragma unroll
                for (int q = 0; q < NW; q++) x[r][q] = src[q];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < W; k++) {
        const uint32_t P = pack_plane(x[0][k >> 2], x[1][k >> 2], x[2][k >> 2], x[3][k >> 2], k & 3);
        const int64_t A = s.out_values + (int64_t)k * s.n + j0;
        if (m == 128) {
            const unsigned a = (unsigned)(A & 3);
            if (a == 0) {
                ((uint32_t*)(out + A))[lane] = P;
            } else {
                const uint32_t prev = __shfl_up_sync(FULL_MASK, P, 1);
                uint32_t* al = (uint32_t*)(out + A - a);
                if (lane > 0) al[lane] = __funnelshift_r(prev, P, 8 * (4 - a));
                else
                    for (unsigned t = 0; t < 4 - a; t++) out[A + t] = (uint8_t)(P >> (8 * t));
                if (lane == 31)
                    for (unsigned t = 0; t < a; t++) out[A + 128 - a + t] = (uint8_t)(P >> (8 * (4 - a + t)));
            }
        } else {
            for (int t = 0; t < mine; t++) out[A + 4 * lane + t] = (uint8_t)(P >> (8 * t));
        }
    }
}
// varints / has_nulls bytes computed on the host: (offset, length <= 10, bytes)
struct SerSmall {
    int64_t off;
    uint8_t len;
    uint8_t b[10];
};
__global__ void serde_small_kernel(const SerSmall* __restrict__ items, int n, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SerSmall it = items[i];
    for (int k = 0; k < it.len; k++) out[it.off + k] = it.b[k];
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
    std::vector<SerSmall> small;   // varints / flags computed here, written by one kernel
    auto add_small = [&](int64_t off, const uint8_t* b, int len) {
        SerSmall it;
        it.off = off;
        it.len = (uint8_t)len;
        memcpy(it.b, b, (size_t)len);
        small.push_back(it);
    };
    // 128-row chunks per partition (fixed-width fast path)
    std::vector<int32_t> chunk_base((size_t)num_parts + 1, 0);
    for (int p = 0; p < num_parts; p++) chunk_base[(size_t)p + 1] = chunk_base[(size_t)p] + (int32_t)((row_offsets[p + 1] - row_offsets[p] + 127) / 128);
    int64_t pos = 0;
    for (int p = 0; p < num_parts; p++) {
        res.part_offsets[p] = pos;
        int64_t n = row_offsets[p + 1] - row_offsets[p];
        for (int c = 0; c < ncols; c++) segs[c][p] = SerSeg{-1, 0, 0, 0, row_offsets[p], n};
        if (n == 0) continue;   // empty partitions write nothing (ipc_compression.rs:68-70)
        uint8_t hdr[10];
        int hl = put_varint((uint64_t)n, hdr);
        add_small(pos, hdr, hl);
        pos += hl;
        for (int c = 0; c < ncols; c++) {
            const Column& col = *b.cols[c];
            SerSeg& s = segs[c][p];
            if (col.type.id == T_NULL) continue;
            bool has_nulls = col.may_have_nulls();
            const uint8_t flag = has_nulls ? 1 : 0;
            add_small(pos, &flag, 1);
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
    if (!small.empty()) {
        Buf dsmall = to_device(ctx, small.data(), small.size() * sizeof(SerSmall));
        serde_small_kernel<<<(unsigned)((small.size() + 255) / 256), 256, 0, ctx.stream>>>(P<SerSmall>(dsmall), (int)small.size(), out);
        LAUNCH_CHECK(ctx);
    }
    Buf dchunk = to_device(ctx, chunk_base.data(), chunk_base.size() * 4);
    const int total_chunks = chunk_base.back();
    int64_t n_rows = b.num_rows;
    for (int c = 0; c < ncols; c++) {
        const Column& col = *b.cols[c];
        if (col.type.id == T_NULL || n_rows == 0) continue;
        Buf dsegs = to_device(ctx, segs[c].data(), segs[c].size() * sizeof(SerSeg));
        const int w = col.type.width();
        if (!col.type.is_varlen() && col.type.id != T_BOOL && (w == 2 || w == 4 || w == 8 || w == 16) && total_chunks > 0 && !getenv("AURON_SERDE_ROWWISE")) {
            const unsigned grid = (unsigned)(((int64_t)total_chunks * 32 + 255) / 256);
            const uint8_t* d = P<uint8_t>(col.data);
            switch (w) {
                case 2: serde_fixed_kernel<2><<<grid, 256, 0, ctx.stream>>>(d, col.vbits(), P<int32_t>(dchunk), num_parts, P<SerSeg>(dsegs), total_chunks, out); break;
                case 4: serde_fixed_kernel<4><<<grid, 256, 0, ctx.stream>>>(d, col.vbits(), P<int32_t>(dchunk), num_parts, P<SerSeg>(dsegs), total_chunks, out); break;
                case 8: serde_fixed_kernel<8><<<grid, 256, 0, ctx.stream>>>(d, col.vbits(), P<int32_t>(dchunk), num_parts, P<SerSeg>(dsegs), total_chunks, out); break;
                default: serde_fixed_kernel<16><<<grid, 256, 0, ctx.stream>>>(d, col.vbits(), P<int32_t>(dchunk), num_parts, P<SerSeg>(dsegs), total_chunks, out); break;
            }
            LAUNCH_CHECK(ctx);
            continue;
        }
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
    }
    ctx.sync();   // (to_device stages host vectors before returning, so one sync at the end is enough)
    return res;
}

// ------------------------------------------------------------------------------------------ read side (IpcReaderExec)
// Inverse of the kernels above (read_batch / read_array, batch_serde.rs:81-101,309-346): the decompressed payload of a
// shuffle segment is a sequence of batches; the host walks the section layout (it decompressed the bytes, so it has them)
// and hands one descriptor per (batch, column) to the device, which rebuilds Arrow columns for ALL batches of the chunk
// at once: byte planes -> values (same warp-per-128-rows scheme, planes read as unaligned 32-bit words), validity / bool
// bits re-packed at the output row offset, utf8 lengths -> offsets by one scan, payload bytes by cooperative copies.
__device__ __forceinline__ uint32_t ld32_any(const uint8_t* p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}
template <int W>
__global__ void __launch_bounds__(256) deser_fixed_kernel(const uint8_t* __restrict__ payload, const DeserSeg* __restrict__ segs, const int32_t* __restrict__ chunk_base,
                                                          int n_segs, int total_chunks, uint8_t* __restrict__ out) {
    const int chunk = (int)(((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5);
    if (chunk >= total_chunks) return;
    const unsigned lane = threadIdx.x & 31;
    int lo = 0, hi = n_segs;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (chunk_base[mid] <= chunk) lo = mid;
        else hi = mid;
    }
    const DeserSeg s = segs[lo];
    const int64_t j0 = (int64_t)(chunk - chunk_base[lo]) * 128 + 4 * lane;
    const int mine = (int)max((int64_t)0, min((int64_t)4, s.n - j0));
    if (mine <= 0) return;
    constexpr int NW = W >= 4 ? W / 4 : 1;
    uint32_t P[W];
#pragma unroll
    for (int k = 0; k < W; k++) {
        const uint8_t* q = payload + s.values_off + (int64_t)k * s.n + j0;
        if (mine == 4) P[k] = ld32_any(q);   // the payload buffer is padded, the funnel's second word stays inside it
        else {
            P[k] = 0;
            for (int t = 0; t < mine; t++) P[k] |= (uint32_t)q[t] << (8 * t);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        if (r >= mine) break;
        const int64_t row = s.out_row0 + j0 + r;
        if constexpr (W == 1) out[row] = (uint8_t)(P[0] >> (8 * r));
        else if constexpr (W == 2) ((uint16_t*)out)[row] = (uint16_t)((P[0] >> (8 * r)) & 0xff) | (uint16_t)(((P[1] >> (8 * r)) & 0xff) << 8);
        else {
            uint32_t* d = (uint32_t*)out + row * NW;
#pragma unroll
            for (int q = 0; q < NW; q++) d[q] = pack_plane(P[4 * q], P[4 * q + 1], P[4 * q + 2], P[4 * q + 3], r);
        }
    }
}
// validity sections / bool values: 32 source bits per thread, OR-ed into the zero-initialised output bitmap at the batch's
// row offset.  bits_off < 0: the batch has no validity section (all rows valid).
__global__ void __launch_bounds__(256) deser_bits_kernel(const uint8_t* __restrict__ payload, const DeserSeg* __restrict__ segs, const int32_t* __restrict__ word_base,
                                                         int n_segs, int total_words, int use_values, uint32_t* __restrict__ out) {
    const int w = blockIdx.x * 256 + threadIdx.x;
    if (w >= total_words) return;
    int lo = 0, hi = n_segs;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (word_base[mid] <= w) lo = mid;
        else hi = mid;
    }
    const DeserSeg s = segs[lo];
    const int64_t j = (int64_t)(w - word_base[lo]) * 32;
    const int cnt = (int)min((int64_t)32, s.n - j);
    if (cnt <= 0) return;
    const int64_t src = use_values ? s.values_off : s.validity_off;
    uint32_t bits = 0xffffffffu;
    if (src >= 0) {
        bits = 0;
        const uint8_t* q = payload + src + (j >> 3);
        for (int t = 0; t < (cnt + 7) / 8; t++) bits |= (uint32_t)q[t] << (8 * t);
    }
    if (cnt < 32) bits &= (1u << cnt) - 1u;
    if (!bits) return;
    const int64_t o = s.out_row0 + j;
    const int sh = (int)(o & 31);
    atomicOr(&out[o >> 5], bits << sh);
    if (sh && (bits >> (32 - sh))) atomicOr(&out[(o >> 5) + 1], bits >> (32 - sh));
}
// utf8 / binary: row lengths from the four transposed length planes
__global__ void __launch_bounds__(256) deser_lengths_kernel(const uint8_t* __restrict__ payload, const DeserSeg* __restrict__ segs, const int32_t* __restrict__ chunk_base,
                                                            int n_segs, int total_chunks, int32_t* __restrict__ lens) {
    const int64_t gi = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int chunk = (int)(gi >> 7);   // 128 rows per chunk, one row per thread
    if (chunk >= total_chunks) return;
    int lo = 0, hi = n_segs;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (chunk_base[mid] <= chunk) lo = mid;
        else hi = mid;
    }
    const DeserSeg s = segs[lo];
    const int64_t j = (int64_t)(chunk - chunk_base[lo]) * 128 + (gi & 127);
    if (j >= s.n) return;
    const uint8_t* q = payload + s.values_off + j;
    lens[s.out_row0 + j] = (int32_t)((uint32_t)q[0] | ((uint32_t)q[s.n] << 8) | ((uint32_t)q[2 * s.n] << 16) | ((uint32_t)q[3 * s.n] << 24));
}
__global__ void __launch_bounds__(128) deser_bytes_kernel(const uint8_t* __restrict__ payload, const DeserCopy* __restrict__ copies, int n, uint8_t* __restrict__ out) {
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= n) return;
    const DeserCopy cp = copies[c];
    warp_copy(out + cp.dst, payload + cp.src, cp.len, threadIdx.x & 31);
}

// Layout walk on the device (used when the payload was decompressed on the GPU and never visits the host): one warp per codec
// stream parses its batches -- varint row count, per column the has_nulls varint and the section sizes, string byte totals by a
// warp reduction over the four length planes.  Pass 1 (segs == nullptr) counts batches, pass 2 fills one DeserSeg per (batch,
// column) with out_row0 = 0 (the host adds the row offsets) and the string byte totals.  flags[stream] != 0: malformed payload
// or a batch that continues in the next stream (the host-side walk handles that case).
__global__ void __launch_bounds__(128) deser_layout_kernel(const uint8_t* __restrict__ payload, const LayoutStream* __restrict__ streams, int n_streams,
                                                           LayoutSchema sch, const int32_t* __restrict__ seg_base, DeserSeg* __restrict__ segs,
                                                           int64_t* __restrict__ sbytes, int64_t* __restrict__ batch_rows, int32_t* __restrict__ counts,
                                                           int32_t* __restrict__ flags) {
    const int si = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (si >= n_streams) return;
    const unsigned lane = threadIdx.x & 31;
    const int64_t end = streams[si].end;
    int64_t pos = streams[si].begin;
    int nb = 0;
    bool bad = false;
    auto varint = [&](uint64_t* v) {
        uint64_t x = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (pos >= end) return false;
            const uint8_t b = payload[pos++];
            x |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) {
                *v = x;
                return true;
            }
        }
        return false;
    };
    while (pos < end && !bad) {
        uint64_t n64;
        if (!varint(&n64) || n64 > 0x7fffffffull) { bad = true; break; }
        const int64_t n = (int64_t)n64;
        const int64_t slot = segs ? (int64_t)seg_base[si] + nb : 0;
        if (segs && lane == 0) batch_rows[slot] = n;
        for (int c = 0; c < sch.ncols && !bad; c++) {
            const int kind = sch.kind[c];
            if (kind == 0) continue;
            uint64_t hn;
            if (!varint(&hn)) { bad = true; break; }
            DeserSeg sg{-1, 0, 0, n};
            if (hn) {
                sg.validity_off = pos;
                pos += (n + 7) / 8;
            }
            sg.values_off = pos;
            int64_t sum = 0;
            if (kind == 1) pos += (n + 7) / 8;
            else if (kind == 2) pos += (int64_t)sch.width[c] * n;
            else {
                if (pos + 4 * n > end) { bad = true; break; }
                const uint8_t* p0 = payload + pos;
                for (int64_t i = lane; i < n; i += 32) sum += (int64_t)((uint32_t)p0[i] | ((uint32_t)p0[n + i] << 8) | ((uint32_t)p0[2 * n + i] << 16) | ((uint32_t)p0[3 * n + i] << 24));
                for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(FULL_MASK, sum, d);
                pos += 4 * n + sum;
            }
            if (pos > end) { bad = true; break; }
            if (segs && lane == 0) {
                segs[slot * sch.ncols + c] = sg;
                sbytes[slot * sch.ncols + c] = sum;
            }
        }
        nb++;
    }
    if (lane == 0) {
        counts[si] = nb;
        flags[si] = bad ? 1 : 0;
    }
}
void deserialize_layout(Ctx& ctx, const uint8_t* dev_payload, const LayoutStream* dev_streams, int n_streams, const LayoutSchema& sch, const int32_t* dev_seg_base,
                        DeserSeg* dev_segs, int64_t* dev_sbytes, int64_t* dev_batch_rows, int32_t* dev_counts, int32_t* dev_flags) {
    if (n_streams <= 0) return;
    deser_layout_kernel<<<(n_streams + 3) / 4, 128, 0, ctx.stream>>>(dev_payload, dev_streams, n_streams, sch, dev_seg_base, dev_segs, dev_sbytes, dev_batch_rows,
                                                                     dev_counts, dev_flags);
    LAUNCH_CHECK(ctx);
}

ColumnPtr deserialize_column(Ctx& ctx, const DType& type, const uint8_t* dev_payload, const std::vector<DeserSeg>& segs, int64_t total_rows,
                             const std::vector<DeserCopy>& byte_copies, int64_t total_bytes) {
    auto col = std::make_shared<Column>();
    col->type = type;
    col->len = total_rows;
    if (type.id == T_NULL) {
        col->null_count = total_rows;
        return col;
    }
    const int n_segs = (int)segs.size();
    bool any_nulls = false;
    std::vector<int32_t> chunk_base((size_t)n_segs + 1, 0), word_base((size_t)n_segs + 1, 0);
    for (int i = 0; i < n_segs; i++) {
        any_nulls = any_nulls || segs[(size_t)i].validity_off >= 0;
        chunk_base[(size_t)i + 1] = chunk_base[(size_t)i] + (int32_t)((segs[(size_t)i].n + 127) / 128);
        word_base[(size_t)i + 1] = word_base[(size_t)i] + (int32_t)((segs[(size_t)i].n + 31) / 32);
    }
    const int total_chunks = chunk_base.back(), total_words = word_base.back();
    Buf dsegs = to_device(ctx, segs.empty() ? (const void*)"" : (const void*)segs.data(), segs.size() * sizeof(DeserSeg));
    Buf dchunks = to_device(ctx, chunk_base.data(), chunk_base.size() * 4);
    Buf dwords = to_device(ctx, word_base.data(), word_base.size() * 4);
    if (any_nulls && total_rows > 0) {
        col->validity = dalloc_zero(ctx, bitmap_alloc_bytes(total_rows));
        col->null_count = -1;
        deser_bits_kernel<<<(unsigned)((total_words + 255) / 256), 256, 0, ctx.stream>>>(dev_payload, P<DeserSeg>(dsegs), P<int32_t>(dwords), n_segs, total_words, 0,
                                                                                         P<uint32_t>(col->validity));
        LAUNCH_CHECK(ctx);
    }
    if (total_rows == 0) {
        if (type.is_varlen()) {
            col->offsets = dalloc_zero(ctx, 4);
            col->data = dalloc(ctx, 1);
        } else col->data = dalloc(ctx, 16);
        return col;
    }
    const unsigned cgrid = (unsigned)(((int64_t)total_chunks * 32 + 255) / 256);
    if (type.id == T_BOOL) {
        col->data = dalloc_zero(ctx, bitmap_alloc_bytes(total_rows));
        deser_bits_kernel<<<(unsigned)((total_words + 255) / 256), 256, 0, ctx.stream>>>(dev_payload, P<DeserSeg>(dsegs), P<int32_t>(dwords), n_segs, total_words, 1,
                                                                                         P<uint32_t>(col->data));
        LAUNCH_CHECK(ctx);
    } else if (type.is_varlen()) {
        AURON_CHECK(total_bytes < (int64_t)INT32_MAX, "shuffle read: more than 2 GiB of string data in one batch");
        Buf lens = dalloc(ctx, (size_t)(total_rows + 1) * 4);
        deser_lengths_kernel<<<(unsigned)(((int64_t)total_chunks * 128 + 255) / 256), 256, 0, ctx.stream>>>(dev_payload, P<DeserSeg>(dsegs), P<int32_t>(dchunks), n_segs,
                                                                                                          total_chunks, P<int32_t>(lens));
        LAUNCH_CHECK(ctx);
        CUDA_OK(cudaMemsetAsync(P<int32_t>(lens) + total_rows, 0, 4, ctx.stream));
        col->offsets = dalloc(ctx, (size_t)(total_rows + 1) * 4);
        Buf tot = dalloc(ctx, 4);
        exclusive_scan_i32(ctx, P<int32_t>(lens), P<int32_t>(col->offsets), total_rows + 1, P<int32_t>(tot));
        col->data = dalloc(ctx, (size_t)std::max<int64_t>(total_bytes, 1));
        col->data_bytes = total_bytes;
        if (!byte_copies.empty()) {
            Buf dcp = to_device(ctx, byte_copies.data(), byte_copies.size() * sizeof(DeserCopy));
            deser_bytes_kernel<<<(unsigned)((byte_copies.size() + 3) / 4), 128, 0, ctx.stream>>>(dev_payload, P<DeserCopy>(dcp), (int)byte_copies.size(), P<uint8_t>(col->data));
            LAUNCH_CHECK(ctx);
        }
    } else {
        const int w = type.width();
        col->data = dalloc(ctx, (size_t)total_rows * w);
        uint8_t* o = P<uint8_t>(col->data);
        switch (w) {
            case 1: deser_fixed_kernel<1><<<cgrid, 256, 0, ctx.stream>>>(dev_payload, P<DeserSeg>(dsegs), P<int32_t>(dchunks), n_segs, total_chunks, o); break;
            case 2: deser_fixed_kernel<2><<<cgrid, 256, 0, ctx.stream>>>(dev_payload, P<DeserSeg>(dsegs), P<int32_t>(dchunks), n_segs, total_chunks, o); break;
            case 4: deser_fixed_kernel<4><<<cgrid, 256, 0, ctx.stream>>>(dev_payload, P<DeserSeg>(dsegs), P<int32_t>(dchunks), n_segs, total_chunks, o); break;
            case 8: deser_fixed_kernel<8><<<cgrid, 256, 0, ctx.stream>>>(dev_payload, P<DeserSeg>(dsegs), P<int32_t>(dchunks), n_segs, total_chunks, o); break;
            case 16: deser_fixed_kernel<16><<<cgrid, 256, 0, ctx.stream>>>(dev_payload, P<DeserSeg>(dsegs), P<int32_t>(dchunks), n_segs, total_chunks, o); break;
            default: fail("shuffle read: unsupported fixed width " + std::to_string(w));
        }
        LAUNCH_CHECK(ctx);
    }
    return col;
}

}  // namespace auron
