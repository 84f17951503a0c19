// k_lz4.cu -- LZ4-frame compression of shuffle blocks on device (row S5 of SURVEY.md section 8a).
//
// The reference compresses every shuffle block on the task's CPU core with lz4_flex's FrameEncoder
// (datafusion-ext-commons/src/io/ipc_compression.rs:35-113,178-197): `block := u32_le compressed_len | LZ4 frame`.
// Leaving that on the host made ShuffleWriterExec 300x slower than every kernel around it (measured: 64M rows x 28 B,
// 1.26 s of host work next to 3 ms of kernels), so the frame is produced on the GPU:
//
//   * an LZ4 frame with independent blocks is a sequence of <= 64 KB blocks that compress independently -> ONE WARP PER
//     64 KB BLOCK, tens of thousands of blocks per chunk;
//   * per step the 32 lanes test 32 consecutive positions against a 4096-entry position table in shared memory (hash of
//     the next 4 bytes); the first lane with a verified candidate wins, the match is extended forward 128 bytes per
//     iteration, and literals / match are emitted with cooperative copies (greedy parse, like LZ4's fast mode);
//   * byte-plane transposed columns (batch_serde.rs:292-305) are long runs of equal bytes in the high planes and noise in
//     the low planes: few, long sequences per block;
//   * a second kernel assembles the partition streams ([u32 len][frame header][block size | data]...[end mark]) at offsets
//     computed on the host from the per-block sizes (one 4-byte D2H per block), blocks that did not shrink are stored
//     raw (high bit of the block size word, as LZ4F does), then ONE D2H moves the finished file image.
//
// Any LZ4-frame decoder reads the result (the tests use Arrow C++'s; the reference reader is lz4_flex's FrameDecoder,
// ipc_compression.rs:115-176).  Roofline: HBM-bound, algorithmic bytes = raw bytes in + compressed bytes out.
#include "device_utils.cuh"
#include "kernels.h"

namespace auron {

#define LAUNCH_CHECK(ctx)            \
    do {                             \
        CUDA_OK(cudaGetLastError()); \
        launch_count(ctx);           \
    } while (0)

constexpr int LZ_HT = 4096;            // position table entries per warp (u16 positions: blocks are <= 64 KB)
constexpr int LZ_MINMATCH = 4, LZ_MFLIMIT = 12, LZ_LASTLITERALS = 5;

__device__ __forceinline__ uint32_t lz_load4(const uint8_t* p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}
// length field continuation: (v - 15) as a run of 255s and one final byte (lane 0 owns the final byte)
__device__ __forceinline__ int lz_emit_ext(uint8_t* dst, int v, unsigned lane) {
    const int r = v - 15, n255 = r / 255;
    for (int i = lane; i < n255; i += 32) dst[i] = 255;
    if (lane == 0) dst[n255] = (uint8_t)(r - n255 * 255);
    return n255 + 1;
}

__global__ void __launch_bounds__(128) lz4_compress_blocks_kernel(const Lz4Block* __restrict__ blocks, int n_blocks, int32_t* __restrict__ out_sizes) {
    __shared__ uint16_t s_ht[4][LZ_HT];
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= n_blocks) return;
    const unsigned lane = threadIdx.x & 31;
    uint16_t* ht = s_ht[threadIdx.x >> 5];
    const Lz4Block blk = blocks[b];
    const uint8_t* __restrict__ src = blk.src;
    uint8_t* __restrict__ dst = blk.dst;
    const int n = blk.len;
    for (int i = lane; i < LZ_HT; i += 32) ht[i] = 0;
    __syncwarp();
    int anchor = 0, ip = 0, op = 0;
    const int mflimit = n - LZ_MFLIMIT, matchlimit = n - LZ_LASTLITERALS;
    while (ip < mflimit) {
        const int q = ip + (int)lane;
        const bool valid = q < mflimit;
        uint32_t seq = 0, h = 0;
        int cand = 0;
        if (valid) {
            seq = lz_load4(src + q);
            h = (seq * 2654435761u) >> 20;
            cand = ht[h];
        }
        __syncwarp();
        if (valid) ht[h] = (uint16_t)q;   // candidates were read before any lane of this step published its position
        // A candidate must agree on 8 bytes, not LZ4's minimum of 4: byte planes with a few distinct values (the third byte
        // of a 204k-valued key) match everywhere for 4-8 bytes, and a warp spends ~400 cycles per sequence -- measured 17.3 ms
        // per 1.8 GB with the 4-byte rule against 10.5 ms with this one, for a file that is 0.8 % larger (638 -> 643 MB).
        const bool hit = valid && cand < q && lz_load4(src + cand) == seq && lz_load4(src + cand + 4) == lz_load4(src + q + 4);
        const unsigned ball = __ballot_sync(0xffffffffu, hit);
        if (ball == 0) {
            ip += 32;
            continue;
        }
        const int l = __ffs(ball) - 1;
        const int m = ip + l;
        const int ref = __shfl_sync(0xffffffffu, cand, l);
        // extend the match forward, 4 bytes per lane per iteration; never past matchlimit
        int mlen = LZ_MINMATCH;
        for (;;) {
            const int j = mlen + (int)lane * 4;
            int eq = 0;   // equal bytes of this lane's 4-byte group
            if (m + j + 4 <= matchlimit) {
                const uint32_t x = lz_load4(src + m + j) ^ lz_load4(src + ref + j);
                eq = x ? ((__ffs(x) - 1) >> 3) : 4;
            } else {
                while (eq < 4 && m + j + eq < matchlimit && src[m + j + eq] == src[ref + j + eq]) eq++;
            }
            const unsigned stop = __ballot_sync(0xffffffffu, eq < 4);
            if (stop == 0) {
                mlen += 128;
                continue;
            }
            const int f = __ffs(stop) - 1;
            mlen += f * 4 + __shfl_sync(0xffffffffu, eq, f);
            break;
        }
        // sequence: token | literal length ext | literals | offset | match length ext
        const int litlen = m - anchor, ml = mlen - LZ_MINMATCH;
        if (lane == 0) dst[op] = (uint8_t)((min(litlen, 15) << 4) | min(ml, 15));
        op += 1;
        if (litlen >= 15) op += lz_emit_ext(dst + op, litlen, lane);
        warp_copy(dst + op, src + anchor, litlen, lane);
        op += litlen;
        if (lane == 0) {
            const int off = m - ref;
            dst[op] = (uint8_t)(off & 255);
            dst[op + 1] = (uint8_t)(off >> 8);
        }
        op += 2;
        if (ml >= 15) op += lz_emit_ext(dst + op, ml, lane);
        ip = m + mlen;
        anchor = ip;
        __syncwarp();
    }
    // last literals
    {
        const int litlen = n - anchor;
        if (lane == 0) dst[op] = (uint8_t)(min(litlen, 15) << 4);
        op += 1;
        if (litlen >= 15) op += lz_emit_ext(dst + op, litlen, lane);
        warp_copy(dst + op, src + anchor, litlen, lane);
        op += litlen;
    }
    if (lane == 0) out_sizes[b] = op;
}

__global__ void __launch_bounds__(128) lz4_assemble_kernel(const Lz4Place* __restrict__ places, int n) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= n) return;
    const unsigned lane = threadIdx.x & 31;
    const Lz4Place pl = places[b];
    uint8_t* d = pl.dst;
    if (pl.flags & 1) {   // first block of a partition stream: u32 frame length + frame header (magic, FLG, BD, HC)
        if (lane < 4) d[-11 + (int)lane] = (uint8_t)(pl.stream_len >> (8 * lane));
        if (lane < 7) d[-7 + (int)lane] = pl.header[lane];
    }
    if (lane < 4) d[lane] = (uint8_t)(pl.size_word >> (8 * lane));
    warp_copy(d + 4, pl.src, pl.len, lane);
    if ((pl.flags & 2) && lane < 4) d[4 + pl.len + lane] = 0;   // end mark
}

// ------------------------------------------------------------------------------------------ decompression (IpcReaderExec)
// One warp per LZ4 block of a frame with independent blocks (what this engine's writer and lz4_flex's FrameEncoder produce):
// the warp walks the sequences -- token, literal length, literals, offset, match length -- with uniform control flow; literals
// are cooperative vector copies from the compressed block, matches are served from a shared-memory ring of the last 4 KB of
// output when they are short and near (the common case in the noisy byte planes), else from the output in global memory
// (long runs of equal bytes in the high planes: pattern copy, byte i = out[pos - offset + i % offset]).
// out_sizes[b] = decoded bytes, or -1 for a malformed block.
constexpr int LZD_RING = 4096;
__global__ void __launch_bounds__(128) lz4_decompress_blocks_kernel(const Lz4DBlock* __restrict__ blocks, int n_blocks, int32_t* __restrict__ out_sizes) {
    __shared__ uint8_t s_ring[4][LZD_RING];
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= n_blocks) return;
    const unsigned lane = threadIdx.x & 31;
    uint8_t* ring = s_ring[threadIdx.x >> 5];
    const Lz4DBlock blk = blocks[b];
    const uint8_t* __restrict__ src = blk.src;
    uint8_t* dst = blk.dst;
    const int n_in = blk.src_len, cap = blk.dst_cap;
    if (blk.stored) {
        if (n_in > cap) {
            if (lane == 0) out_sizes[b] = -1;
            return;
        }
        warp_copy(dst, src, n_in, lane);
        if (lane == 0) out_sizes[b] = n_in;
        return;
    }
    int ip = 0, op = 0, ring_from = 0;
    bool bad = false;
    while (ip < n_in) {
        const uint32_t token = src[ip++];
        int litlen = (int)(token >> 4);
        if (litlen == 15) {
            for (;;) {
                if (ip >= n_in) { bad = true; break; }
                const uint32_t x = src[ip++];
                litlen += (int)x;
                if (x != 255) break;
            }
            if (bad) break;
        }
        if (litlen > n_in - ip || litlen > cap - op) { bad = true; break; }
        if (litlen > 0) {
            if (litlen <= 256) {
                for (int i = lane; i < litlen; i += 32) {
                    const uint8_t c = src[ip + i];
                    dst[op + i] = c;
                    ring[(op + i) & (LZD_RING - 1)] = c;
                }
            } else {
                warp_copy(dst + op, src + ip, litlen, lane);
                ring_from = op + litlen;   // the ring does not hold this literal
            }
            ip += litlen;
            op += litlen;
        }
        if (ip >= n_in) break;   // the last sequence of a block has no match part
        if (ip + 2 > n_in) { bad = true; break; }
        const int off = (int)((uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8));
        ip += 2;
        int mlen = (int)(token & 15);
        if (mlen == 15) {
            for (;;) {
                if (ip >= n_in) { bad = true; break; }
                const uint32_t x = src[ip++];
                mlen += (int)x;
                if (x != 255) break;
            }
            if (bad) break;
        }
        mlen += LZ_MINMATCH;
        if (off == 0 || off > op || mlen > cap - op) { bad = true; break; }
        __syncwarp();   // the referenced bytes were written by other lanes
        const int from = op - off;
        if (mlen <= 256 && from >= ring_from && off <= LZD_RING - 256) {
            for (int i = lane; i < mlen; i += 32) {
                const int k = off >= mlen ? i : (int)((unsigned)i % (unsigned)off);
                const uint8_t c = ring[(from + k) & (LZD_RING - 1)];
                dst[op + i] = c;
                ring[(op + i) & (LZD_RING - 1)] = c;   // never a slot that is still to be read: off + mlen <= LZD_RING
            }
        } else if (off >= mlen) {
            warp_copy(dst + op, dst + from, mlen, lane);   // disjoint ranges
            ring_from = op + mlen;
        } else {
            // overlapping run: every output byte depends only on bytes before `op`
            if (off >= 4 && mlen >= 256) {
                // period >= 4: replicate 4-byte words; lane handles words w = lane, lane + 32, ... of the run
                for (int i = lane * 4; i < mlen; i += 128) {
                    uint32_t v = 0;
                    const int m4 = min(4, mlen - i);
                    for (int t = 0; t < m4; t++) v |= (uint32_t)dst[from + (int)((unsigned)(i + t) % (unsigned)off)] << (8 * t);
                    for (int t = 0; t < m4; t++) dst[op + i + t] = (uint8_t)(v >> (8 * t));
                }
            } else if (off < 4 && mlen >= 64) {
                // tiny period (runs of one byte value dominate the high planes): build the 4-byte pattern word once, store words
                uint32_t pat = 0;
                for (int t = 0; t < 4; t++) pat |= (uint32_t)dst[from + (t % off)] << (8 * t);
                // output byte j (absolute op + j) = pattern byte j % off; rotate the word so that aligned stores line up
                const int head = (int)((4 - ((uintptr_t)(dst + op) & 3)) & 3);
                for (int i = lane; i < min(head, mlen); i += 32) dst[op + i] = dst[from + (i % off)];
                const int body = (mlen - head) >> 2;
                uint32_t w = 0;
                for (int t = 0; t < 4; t++) w |= (uint32_t)((pat >> (8 * ((head + t) % off))) & 0xff) << (8 * t);
                // with off in {1, 2} every word of the body is identical; off == 3 has period 3 words
                uint32_t* dw = (uint32_t*)(dst + op + head);
                if (off == 3) {
                    for (int q = lane; q < body; q += 32) {
                        uint32_t x = 0;
                        for (int t = 0; t < 4; t++) x |= (uint32_t)((pat >> (8 * ((head + 4 * q + t) % 3))) & 0xff) << (8 * t);
                        dw[q] = x;
                    }
                } else {
                    for (int q = lane; q < body; q += 32) dw[q] = w;
                }
                for (int i = head + 4 * body + lane; i < mlen; i += 32) dst[op + i] = dst[from + (i % off)];
            } else {
                for (int i = lane; i < mlen; i += 32) dst[op + i] = dst[from + (int)((unsigned)i % (unsigned)off)];
            }
            ring_from = op + mlen;
        }
        op += mlen;
        __syncwarp();
    }
    if (lane == 0) out_sizes[b] = bad ? -1 : op;
}

void lz4_decompress_blocks(Ctx& ctx, const Lz4DBlock* dev_blocks, int n_blocks, int32_t* dev_sizes) {
    if (n_blocks <= 0) return;
    ProfScope ps(ctx, "lz4_decompress");
    lz4_decompress_blocks_kernel<<<(n_blocks + 3) / 4, 128, 0, ctx.stream>>>(dev_blocks, n_blocks, dev_sizes);
    LAUNCH_CHECK(ctx);
}

void lz4_compress_blocks(Ctx& ctx, const Lz4Block* dev_blocks, int n_blocks, int32_t* dev_sizes) {
    if (n_blocks <= 0) return;
    ProfScope ps(ctx, "lz4_compress");
    lz4_compress_blocks_kernel<<<(n_blocks + 3) / 4, 128, 0, ctx.stream>>>(dev_blocks, n_blocks, dev_sizes);
    LAUNCH_CHECK(ctx);
}
void lz4_assemble(Ctx& ctx, const Lz4Place* dev_places, int n) {
    if (n <= 0) return;
    ProfScope ps(ctx, "lz4_assemble");
    lz4_assemble_kernel<<<(n + 3) / 4, 128, 0, ctx.stream>>>(dev_places, n);
    LAUNCH_CHECK(ctx);
}

}  // namespace auron
