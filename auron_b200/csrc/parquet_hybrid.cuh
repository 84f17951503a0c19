// parquet_hybrid.cuh -- device-side pieces shared by the Parquet decode kernels (k_parquet.cu) and the fused
// scan -> filter -> aggregate kernels (k_fused.cu): unaligned loads, the RLE / bit-packed hybrid stream reader with
// checkpoints, and the warp-parallel parse of definition-level streams.
#pragma once
#include "device_utils.cuh"
#include "parquet_dev.h"

namespace auron {

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

constexpr int PQ_WARPS = 4;
constexpr int PQ_TILE = 1024;   // rows per decode tile (one warp)

// (HybridCk, the checkpoint of a hybrid stream, lives in parquet_dev.h)
struct PqTile {
    int32_t page, row0, n, pad;
    int64_t v0;   // non-null values of the page before this tile
    HybridCk def, idx;
};
__device__ __forceinline__ HybridCk hybrid_save(const Hybrid& h, const uint8_t* base) {
    HybridCk c;
    c.p_off = (int32_t)(h.p - base);
    c.run_remaining = h.run_remaining;
    c.bp_base_off = (int32_t)(h.bp_base - base);
    c.bp_consumed = h.bp_consumed;
    c.rle_value = h.rle_value;
    c.is_rle = h.is_rle ? 1 : 0;
    return c;
}
__device__ __forceinline__ void hybrid_restore(Hybrid& h, const HybridCk& c, const uint8_t* base, const uint8_t* end, int bw) {
    h.p = base + c.p_off;
    h.end = end;
    h.bw = bw;
    h.run_remaining = c.run_remaining;
    h.is_rle = c.is_rle != 0;
    h.rle_value = c.rle_value;
    h.bp_base = base + c.bp_base_off;
    h.bp_consumed = c.bp_consumed;
}
// advance a stream by n values without materialising them (header walk only)
__device__ __forceinline__ void hybrid_skip(Hybrid& h, int n) {
    while (n > 0) {
        if (h.run_remaining == 0) h.next_run();
        int t = min(n, h.run_remaining);
        h.run_remaining -= t;
        if (!h.is_rle) h.bp_consumed += t;
        n -= t;
    }
}
// number of ones among the next m values of a bit-width-1 stream (m <= 1024); warp-cooperative, advances the stream
__device__ __forceinline__ int hybrid_count_ones(Hybrid& h, int m, unsigned lane) {
    int cnt = 0;
    while (m > 0) {
        if (h.run_remaining == 0) h.next_run();
        int t = min(m, h.run_remaining);
        if (h.is_rle) cnt += (h.rle_value & 1) ? t : 0;
        else {
            int64_t s = (int64_t)h.bp_consumed + 32 * (int64_t)lane;   // this lane's first bit
            int mine = min(32, t - 32 * (int)lane);
            int c = 0;
            if (mine > 0) {
                const uint8_t* q = h.bp_base + (s >> 3);
                uint64_t w = 0;
#pragma unroll
                for (int i = 0; i < 5; i++) w |= (uint64_t)q[i] << (8 * i);
                uint32_t bits = (uint32_t)(w >> (s & 7));
                if (mine < 32) bits &= (1u << mine) - 1u;
                c = __popc(bits);
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(FULL_MASK, c, d);
            cnt += c;
            h.bp_consumed += t;
        }
        h.run_remaining -= t;
        m -= t;
    }
    return cnt;
}

// 64-bit window starting at an arbitrary bit position (two aligned 32-bit pairs + funnel shifts)
__device__ __forceinline__ uint32_t extract_bits(const uint8_t* base, int64_t bitpos, int bw) {
    const uint8_t* q = base + (bitpos >> 3);
    uintptr_t a = (uintptr_t)q;
    const uint32_t* wp = (const uint32_t*)(a & ~(uintptr_t)3);
    unsigned sh = (unsigned)(a & 3) * 8 + (unsigned)(bitpos & 7);   // 0..31
    uint32_t v = __funnelshift_r(wp[0], wp[1], sh);   // bits [sh, sh + 32) of the aligned 64-bit pair: enough for any bw <= 32
    return bw == 32 ? v : (v & ((1u << bw) - 1u));
}

// ---- parallel parse of a bit-width-1 hybrid stream (definition levels) -------------------------------------------
// Level streams are chains of tiny runs (a NULL every ~30 rows gives ~60 runs per 1024 rows); walking them serially
// costs a dependent chain per run.  Instead the stream is treated as a finite-state transducer over 1 KiB windows:
//   1. lane L owns bytes [32L, 32L+32) of the window and computes, for EVERY entry offset o in its block, where a
//      parse entering at o leaves the block (dynamic programming from o = 31 down: exit[o] = next(o) >= 32 ? next(o)
//      : exit[next(o)]) -- no knowledge of the true entry point needed;
//   2. lane 0 chains the 32 tables from the window's known entry position: 32 lookups give every lane its real entry;
//   3. lanes count the rows of their own runs, a warp scan turns the counts into row bases;
//   4. lanes emit their runs' bits into the page bitmap concurrently (atomicOr on partial words).
__device__ __forceinline__ uint32_t extract_bits(const uint8_t* base, int64_t bitpos, int bw);
struct LvlHdr {
    int hl;        // header bytes
    int payload;   // payload bytes
    int rows;      // values in the run
    int kind;      // 0 bit-packed, 1 RLE ones, 2 RLE zeros
};
__device__ __forceinline__ LvlHdr lvl_parse(const uint8_t* base, int p, int len) {
    LvlHdr r;
    uint32_t h = base[p];
    int hl = 1;
    if (h & 0x80) {
        h &= 0x7f;
        int sh = 7;
        while (p + hl < len) {
            uint32_t b = base[p + hl];
            hl++;
            h |= (b & 0x7f) << sh;
            sh += 7;
            if (!(b & 0x80) || sh > 28) break;
        }
    }
    r.hl = hl;
    if (h & 1) {
        r.kind = 0;
        r.payload = (int)(h >> 1);
        r.rows = (int)(h >> 1) * 8;
    } else {
        r.payload = 1;
        r.rows = (int)(h >> 1);
        r.kind = (p + hl < len && (base[p + hl] & 1)) ? 1 : 2;
    }
    return r;
}
// set bits [r, r+t) of a zero-initialised bitmap
__device__ __forceinline__ void bm_set_range(uint32_t* bm, int64_t r, int64_t t) {
    while (t > 0) {
        int sh = (int)(r & 31);
        int take = (int)min((int64_t)(32 - sh), t);
        uint32_t bits = take == 32 ? 0xffffffffu : (((1u << take) - 1u) << sh);
        if (take == 32) bm[r >> 5] = bits;
        else atomicOr(&bm[r >> 5], bits);
        r += take;
        t -= take;
    }
}
// copy t bits starting at byte `src` (bit 0) to bitmap position r
__device__ __forceinline__ void bm_copy_bits(uint32_t* bm, int64_t r, const uint8_t* src, int64_t t) {
    int64_t done = 0;
    while (done < t) {
        int sh = (int)((r + done) & 31);
        int take = (int)min((int64_t)(32 - sh), t - done);
        uint32_t bits = extract_bits(src, done, 32);
        if (take < 32) bits &= (1u << take) - 1u;
        if (bits) atomicOr(&bm[(r + done) >> 5], bits << sh);
        done += take;
    }
}
// `bit_off`: position of the page's first row in `bm` (bitmaps shared by several pages: partial words are ORed atomically).
// The caller orders the emitted bits before reading them back (fence, or __syncwarp for a shared-memory bitmap).
__device__ inline void lvl_page_bits(const uint8_t* base, int len, int64_t rows, unsigned lane, uint32_t* bm, int (*exitT)[33], int* s_entry,
                                     int64_t bit_off = 0) {
    int64_t row_base = 0;
    int pos = 0;   // absolute byte offset of the next run header
    while (pos < len && row_base < rows) {
        const int win = pos;
        const int b0 = win + 32 * (int)lane;
        // 1. exit table of this lane's block (relative to the block start; >= 32 means "left the block")
        for (int o = 31; o >= 0; o--) {
            int p = b0 + o, nx;
            if (p >= len) nx = 32;
            else {
                const uint32_t h0 = base[p];
                if (!(h0 & 0x80)) nx = o + ((h0 & 1) ? 1 + (int)(h0 >> 1) : 2);   // one-byte header: bit-packed groups / RLE value byte
                else {
                    LvlHdr h = lvl_parse(base, p, len);
                    nx = o + h.hl + h.payload;
                }
            }
            exitT[lane][o] = nx >= 32 ? nx : exitT[lane][nx];
        }
        __syncwarp();
        // 2. chain
        if (lane == 0) {
            int rel = 0;
            for (int L = 0; L < 32; L++) {
                if (rel >= 32 * L && rel < 32 * (L + 1)) {
                    s_entry[L] = rel - 32 * L;
                    rel = 32 * L + exitT[L][rel - 32 * L];
                } else s_entry[L] = -1;
            }
            s_entry[32] = rel;
        }
        __syncwarp();
        // 3. rows of my runs
        const int e = s_entry[lane];
        int64_t cnt = 0;
        if (e >= 0) {
            int o = e;
            while (o < 32 && b0 + o < len) {
                LvlHdr h = lvl_parse(base, b0 + o, len);
                cnt += h.rows;
                o += h.hl + h.payload;
            }
        }
        int64_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int64_t t = __shfl_up_sync(FULL_MASK, inc, d);
            if ((int)lane >= d) inc += t;
        }
        const int64_t total = __shfl_sync(FULL_MASK, inc, 31);
        // 4. emit: short runs by their owner lane; long RLE runs of ones (a page without NULLs is ONE such run of up to 20,000
        //    rows -- a single lane needed ~600 serial word stores for it) are handed to the whole warp, 32 words per step
        int64_t long_r[4], long_t[4];
        int n_long = 0;
        if (e >= 0) {
            int o = e;
            int64_t r = row_base + inc - cnt;
            while (o < 32 && b0 + o < len && r < rows) {
                LvlHdr h = lvl_parse(base, b0 + o, len);
                int64_t t = min((int64_t)h.rows, rows - r);
                if (h.kind == 1) {
                    if (t > 128 && n_long < 4) {
                        long_r[n_long] = r + bit_off;
                        long_t[n_long] = t;
                        n_long++;
                    } else bm_set_range(bm, r + bit_off, t);
                } else if (h.kind == 0) bm_copy_bits(bm, r + bit_off, base + b0 + o + h.hl, t);
                r += h.rows;
                o += h.hl + h.payload;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned owners = __ballot_sync(FULL_MASK, k < n_long);
            while (owners) {
                const int l = __ffs(owners) - 1;
                owners &= owners - 1;
                const int64_t r = __shfl_sync(FULL_MASK, long_r[k], l), t = __shfl_sync(FULL_MASK, long_t[k], l);
                const int64_t w0 = r >> 5, w1 = (r + t - 1) >> 5;   // t > 128: at least three words apart
                if (lane == 0) atomicOr(&bm[w0], 0xffffffffu << (r & 31));
                if (lane == 1) atomicOr(&bm[w1], 0xffffffffu >> (31 - ((r + t - 1) & 31)));
                for (int64_t w = w0 + 1 + lane; w < w1; w += 32) bm[w] = 0xffffffffu;
            }
        }
        row_base += total;
        pos = win + s_entry[32];
        __syncwarp();
    }
}

}  // namespace auron
