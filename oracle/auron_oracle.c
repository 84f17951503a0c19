/*
 * auron_oracle.c -- CPU restatement (plain C) of the reference algorithms on the
 * hot path.  TEST INFRASTRUCTURE ONLY: nothing in the product path (auron_b200/,
 * libauron_b200.so) may link, import or call this file.  Only tests/, the smoke
 * check in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py use it, and only as the checker / the timed CPU arm.
 *
 * Parity status: the hash functions, varint, batch serde, casts and partitioner
 * are pinned against the golden vectors held by the reference's own unit tests
 * (see tests/test_oracle_golden.py for the vectors and their file:line).  The
 * reference engine itself (Rust nightly + git-patched arrow-rs/datafusion)
 * cannot be compiled in this image (no cargo/rustc), so oracle/_ref is absent.
 *
 * Every function cites the reference file:line it restates.  Paths are relative
 * to native-engine/ in apache/auron.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* murmur3_x86_32, Spark flavour                                       */
/* datafusion-ext-commons/src/hash/mur.rs:19-87                         */
/* ------------------------------------------------------------------ */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t mix_k1(uint32_t k1) {            /* mur.rs:38-43 */
    k1 *= 0xcc9e2d51u; k1 = rotl32(k1, 15); k1 *= 0x1b873593u; return k1;
}
static inline uint32_t mix_h1(uint32_t h1, uint32_t k1) { /* mur.rs:46-51 */
    h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5u + 0xe6546b64u; return h1;
}
static inline uint32_t fmix(uint32_t h1, uint32_t len) { /* mur.rs:54-62 */
    h1 ^= len; h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
    return h1;
}
/* mur.rs:19-30: whole 4-byte words little-endian, then every trailing byte is
 * sign-extended and mixed as its own block (Spark's hashUnsafeBytes quirk). */
API int32_t orc_murmur3_bytes(const uint8_t* data, int64_t len, int32_t seed) {
    uint32_t h1 = (uint32_t)seed;
    int64_t aligned = len - len % 4;
    for (int64_t i = 0; i < aligned; i += 4) {
        uint32_t w; memcpy(&w, data + i, 4);
        h1 = mix_h1(h1, mix_k1(w));
    }
    for (int64_t i = aligned; i < len; i++) {
        int32_t half = (int32_t)(int8_t)data[i];
        h1 = mix_h1(h1, mix_k1((uint32_t)half));
    }
    return (int32_t)fmix(h1, (uint32_t)len);
}

/* ------------------------------------------------------------------ */
/* xxhash64                                                            */
/* datafusion-ext-commons/src/hash/xxhash.rs:30-120                     */
/* ------------------------------------------------------------------ */
#define P64_1 0x9E3779B185EBCA87ull
#define P64_2 0xC2B2AE3D27D4EB4Full
#define P64_3 0x165667B19E3779F9ull
#define P64_4 0x85EBCA77C2B2AE63ull
#define P64_5 0x27D4EB2F165667C5ull
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t xxr(uint64_t acc, uint64_t in) { acc += in * P64_2; acc = rotl64(acc, 31); acc *= P64_1; return acc; }
static inline uint64_t xxm(uint64_t h, uint64_t acc) { h ^= xxr(0, acc); h *= P64_1; h += P64_4; return h; }
API int64_t orc_xxhash64_bytes(const uint8_t* in, int64_t len, int64_t seed_) {
    uint64_t seed = (uint64_t)seed_, h; int64_t rem = len, off = 0;
    if (rem >= 32) {
        uint64_t a1 = seed + P64_1 + P64_2, a2 = seed + P64_2, a3 = seed, a4 = seed - P64_1;
        while (rem >= 32) {
            uint64_t w;
            memcpy(&w, in + off, 8); a1 = xxr(a1, w); off += 8;
            memcpy(&w, in + off, 8); a2 = xxr(a2, w); off += 8;
            memcpy(&w, in + off, 8); a3 = xxr(a3, w); off += 8;
            memcpy(&w, in + off, 8); a4 = xxr(a4, w); off += 8;
            rem -= 32;
        }
        h = rotl64(a1, 1) + rotl64(a2, 7) + rotl64(a3, 12) + rotl64(a4, 18);
        h = xxm(h, a1); h = xxm(h, a2); h = xxm(h, a3); h = xxm(h, a4);
    } else {
        h = seed + P64_5;
    }
    h += (uint64_t)len;
    while (rem >= 8) {
        uint64_t w; memcpy(&w, in + off, 8);
        h ^= xxr(0, w); h = rotl64(h, 27); h *= P64_1; h += P64_4; off += 8; rem -= 8;
    }
    if (rem >= 4) {
        uint32_t w; memcpy(&w, in + off, 4);
        h ^= (uint64_t)w * P64_1; h = rotl64(h, 23); h *= P64_2; h += P64_3; off += 4; rem -= 4;
    }
    while (rem) { h ^= (uint64_t)in[off] * P64_5; h = rotl64(h, 11); h *= P64_1; off++; rem--; }
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32;
    return (int64_t)h;
}

/* ------------------------------------------------------------------ */
/* column hashing: datafusion-ext-commons/src/spark_hash.rs:46-224      */
/* hashes[] is pre-seeded by the caller (create_hashes :52-57); a NULL  */
/* leaves the running hash unchanged (:78-84).                          */
/* kind: 0 murmur3 (hashes32), 1 xxhash64 (hashes64)                    */
/* ------------------------------------------------------------------ */
static inline int bit_get(const uint8_t* bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

/* fixed-width values hashed as `width_hashed` LE bytes of the sign-extended value
 * (i8/i16/i32/date32 -> i32 :160-168 ; i64/ts/date64 -> i64 ; f32/f64 raw bits) */
API void orc_hash_fixed(int kind, const uint8_t* values, int width_in, int width_hashed, int is_signed,
                        const uint8_t* validity, int64_t n, int32_t* h32, int64_t* h64) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        if (validity && !bit_get(validity, i)) continue;
        uint8_t buf[16];
        if (width_in == width_hashed) {
            memcpy(buf, values + i * width_in, width_in);
        } else { /* widen small ints to i32 */
            int64_t v = 0;
            if (width_in == 1) v = is_signed ? (int64_t)*(const int8_t*)(values + i) : (int64_t)values[i];
            else if (width_in == 2) { int16_t t; memcpy(&t, values + i * 2, 2); v = is_signed ? (int64_t)t : (int64_t)(uint16_t)t; }
            int32_t v32 = (int32_t)v; memcpy(buf, &v32, 4);
        }
        if (kind == 0) h32[i] = orc_murmur3_bytes(buf, width_hashed, h32[i]);
        else h64[i] = orc_xxhash64_bytes(buf, width_hashed, h64[i]);
    }
}
/* bool -> u32 1/0 (spark_hash.rs:131-158) */
API void orc_hash_bool(int kind, const uint8_t* bits, const uint8_t* validity, int64_t n, int32_t* h32, int64_t* h64) {
    for (int64_t i = 0; i < n; i++) {
        if (validity && !bit_get(validity, i)) continue;
        uint32_t v = (uint32_t)bit_get(bits, i);
        if (kind == 0) h32[i] = orc_murmur3_bytes((const uint8_t*)&v, 4, h32[i]);
        else h64[i] = orc_xxhash64_bytes((const uint8_t*)&v, 4, h64[i]);
    }
}
/* utf8 / binary: raw bytes (spark_hash.rs:196-207) */
API void orc_hash_bytes(int kind, const int32_t* offsets, const uint8_t* data, const uint8_t* validity,
                        int64_t n, int32_t* h32, int64_t* h64) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        if (validity && !bit_get(validity, i)) continue;
        const uint8_t* p = data + offsets[i]; int64_t len = offsets[i + 1] - offsets[i];
        if (kind == 0) h32[i] = orc_murmur3_bytes(p, len, h32[i]);
        else h64[i] = orc_xxhash64_bytes(p, len, h64[i]);
    }
}
/* partition id = rem_euclid(hash, N)   datafusion-ext-plans/src/shuffle/mod.rs:178-188 */
API void orc_pmod(const int32_t* h32, int64_t n, int32_t num_parts, int32_t* out) {
    for (int64_t i = 0; i < n; i++) {
        int32_t r = h32[i] % num_parts; if (r < 0) r += num_parts; out[i] = r;
    }
}

/* ------------------------------------------------------------------ */
/* varint: datafusion-ext-commons/src/io/mod.rs:61-84                   */
/* ------------------------------------------------------------------ */
API int64_t orc_write_len(uint64_t len, uint8_t* out) {
    int64_t n = 0;
    while (len >= 128) { out[n++] = (uint8_t)(128 + (len % 128)); len /= 128; }
    out[n++] = (uint8_t)len;
    return n;
}
API int64_t orc_read_len(const uint8_t* in, uint64_t* len) {
    uint64_t v = 0, factor = 1; int64_t n = 0;
    for (;;) {
        uint8_t b = in[n++];
        if (b < 128) { v += (uint64_t)b * factor; break; }
        v += (uint64_t)(b - 128) * factor; factor *= 128;
    }
    *len = v; return n;
}

/* ------------------------------------------------------------------ */
/* batch serde: datafusion-ext-commons/src/io/batch_serde.rs            */
/* primitive :273-307, bool :557-577, utf8 :603-633, bits :190-217      */
/* byte-plane transpose: transpose(w x n) :292-305                      */
/* ------------------------------------------------------------------ */
static int64_t write_bits(const uint8_t* bits, int64_t bit_off, int64_t n, uint8_t* out) {
    int64_t nb = (n + 7) / 8;
    memset(out, 0, nb);
    for (int64_t i = 0; i < n; i++) if (bit_get(bits, bit_off + i)) out[i >> 3] |= (uint8_t)(1u << (i & 7));
    return nb;
}
/* returns bytes written */
API int64_t orc_serde_write_fixed(const uint8_t* values, int width, const uint8_t* validity, int64_t bit_off,
                                  int64_t n, uint8_t* out) {
    int64_t p = 0;
    if (validity) { p += orc_write_len(1, out + p); p += write_bits(validity, bit_off, n, out + p); }
    else p += orc_write_len(0, out + p);
    if (width > 1) {
        for (int b = 0; b < width; b++)
            for (int64_t i = 0; i < n; i++) out[p + (int64_t)b * n + i] = values[i * width + b];
    } else memcpy(out + p, values, n);
    return p + (int64_t)width * n;
}
API int64_t orc_serde_write_bool(const uint8_t* bits, int64_t val_bit_off, const uint8_t* validity, int64_t bit_off,
                                 int64_t n, uint8_t* out) {
    int64_t p = 0;
    if (validity) { p += orc_write_len(1, out + p); p += write_bits(validity, bit_off, n, out + p); }
    else p += orc_write_len(0, out + p);
    p += write_bits(bits, val_bit_off, n, out + p);
    return p;
}
API int64_t orc_serde_write_bytes(const int32_t* offsets, const uint8_t* data, const uint8_t* validity, int64_t bit_off,
                                  int64_t n, uint8_t* out) {
    int64_t p = 0;
    if (validity) { p += orc_write_len(1, out + p); p += write_bits(validity, bit_off, n, out + p); }
    else p += orc_write_len(0, out + p);
    for (int b = 0; b < 4; b++)
        for (int64_t i = 0; i < n; i++) {
            int32_t len = offsets[i + 1] - offsets[i];
            out[p + (int64_t)b * n + i] = (uint8_t)((uint32_t)len >> (8 * b));
        }
    p += 4 * n;
    int64_t nbytes = (int64_t)offsets[n] - offsets[0];
    memcpy(out + p, data + offsets[0], nbytes);
    return p + nbytes;
}
/* readers: return bytes consumed; has_nulls out; validity (ceil(n/8)) out */
API int64_t orc_serde_read_fixed(const uint8_t* in, int width, int64_t n, uint8_t* values, uint8_t* validity, int* has_nulls) {
    uint64_t hn; int64_t p = orc_read_len(in, &hn); *has_nulls = (int)hn;
    if (hn == 1) { memcpy(validity, in + p, (n + 7) / 8); p += (n + 7) / 8; }
    if (width > 1) {
        for (int b = 0; b < width; b++)
            for (int64_t i = 0; i < n; i++) values[i * width + b] = in[p + (int64_t)b * n + i];
    } else memcpy(values, in + p, n);
    return p + (int64_t)width * n;
}
API int64_t orc_serde_read_bool(const uint8_t* in, int64_t n, uint8_t* bits, uint8_t* validity, int* has_nulls) {
    uint64_t hn; int64_t p = orc_read_len(in, &hn); *has_nulls = (int)hn;
    if (hn == 1) { memcpy(validity, in + p, (n + 7) / 8); p += (n + 7) / 8; }
    memcpy(bits, in + p, (n + 7) / 8);
    return p + (n + 7) / 8;
}
/* offsets must hold n+1 entries; data_out may be NULL to size first (returns consumed; *data_len set) */
API int64_t orc_serde_read_bytes(const uint8_t* in, int64_t n, int32_t* offsets, uint8_t* data_out, int64_t* data_len,
                                 uint8_t* validity, int* has_nulls) {
    uint64_t hn; int64_t p = orc_read_len(in, &hn); *has_nulls = (int)hn;
    if (hn == 1) { memcpy(validity, in + p, (n + 7) / 8); p += (n + 7) / 8; }
    int32_t cur = 0;
    for (int64_t i = 0; i < n; i++) {
        uint32_t len = 0;
        for (int b = 0; b < 4; b++) len |= (uint32_t)in[p + (int64_t)b * n + i] << (8 * b);
        offsets[i] = cur; cur += (int32_t)len;
    }
    offsets[n] = cur; p += 4 * n; *data_len = cur;
    if (data_out) memcpy(data_out, in + p, cur);
    return p + cur;
}

/* ------------------------------------------------------------------ */
/* casts: datafusion-ext-commons/src/arrow/cast.rs                      */
/* ------------------------------------------------------------------ */
/* to_integer :394-468; bits = 8/16/32/64; returns 1 if valid */
API int orc_str_to_int(const uint8_t* s, int64_t len, int bits, int64_t* out) {
    if (len == 0) return 0;
    int64_t minv = (bits == 64) ? INT64_MIN : -((int64_t)1 << (bits - 1));
    int negative = s[0] == '-'; int64_t off = 0;
    if (negative || s[0] == '+') { off = 1; if (len == 1) return 0; }
    int64_t stop = minv / 10, result = 0;
    while (off < len) {
        uint8_t b = s[off++];
        if (b == '.') break;
        if (b < '0' || b > '9') return 0;
        if (result < stop) return 0;
        /* result*10 - digit in the target width, wrapping like the Rust release build */
        int64_t r;
        if (bits == 64) r = (int64_t)((uint64_t)result * 10ull - (uint64_t)(b - '0'));
        else if (bits == 32) r = (int32_t)((uint32_t)result * 10u - (uint32_t)(b - '0'));
        else if (bits == 16) r = (int16_t)((uint16_t)((uint16_t)result * 10u - (uint16_t)(b - '0')));
        else r = (int8_t)((uint8_t)((uint8_t)result * 10u - (uint8_t)(b - '0')));
        result = r;
        if (result > 0) return 0;
    }
    while (off < len) { if (s[off] < '0' || s[off] > '9') return 0; off++; }
    if (!negative) {
        if (result == minv) return 0;       /* -min overflows -> negative -> None */
        result = -result;
        if (result < 0) return 0;
    }
    *out = result; return 1;
}
static int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
    y -= m <= 2; int64_t era = (y >= 0 ? y : y - 399) / 400; unsigned yoe = (unsigned)(y - era * 400);
    unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1; unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int64_t)doe - 719468;
}
static int is_leap(int64_t y) { return (y % 4 == 0 && y % 100 != 0) || y % 400 == 0; }
static int valid_digits(int seg, int digits) { return (seg == 0 && digits >= 4 && digits <= 7) || (seg != 0 && digits > 0 && digits <= 2); }
/* to_date :471-529 */
API int orc_str_to_date(const uint8_t* s0, int64_t len0, int32_t* out) {
    /* Rust str::trim(): unicode whitespace; ASCII subset suffices for the vectors */
    int64_t a = 0, e = len0;
    while (a < e && (s0[a] == ' ' || (s0[a] >= 9 && s0[a] <= 13))) a++;
    while (e > a && (s0[e - 1] == ' ' || (s0[e - 1] >= 9 && s0[e - 1] <= 13))) e--;
    const uint8_t* s = s0 + a; int64_t len = e - a;
    if (len == 0) return 0;
    int seg[3] = {1, 1, 1}; int sign = 1, i = 0, cur = 0, digits = 0; int64_t j = 0;
    if (s[j] == '-' || s[j] == '+') { sign = s[j] == '-' ? -1 : 1; j++; }
    while (j < len && (i < 3 && !(s[j] == ' ' || s[j] == 'T'))) {
        uint8_t b = s[j];
        if (i < 2 && b == '-') {
            if (!valid_digits(i, digits)) return 0;
            seg[i] = cur; cur = 0; digits = 0; i++;
        } else {
            int pv = (int)b - '0';
            if (pv < 0 || pv > 9) return 0;
            cur = cur * 10 + pv; digits++;
        }
        j++;
    }
    if (!valid_digits(i, digits)) return 0;
    if (i < 2 && j < len) return 0;
    seg[i] = cur;
    if (seg[0] > 9999 || seg[1] > 12 || seg[2] > 31) return 0;
    int64_t y = (int64_t)sign * seg[0]; int m = seg[1], d = seg[2];
    if (m < 1 || d < 1) return 0;
    static const int mdays[] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    int md = mdays[m - 1] + (m == 2 && is_leap(y));
    if (d > md) return 0;
    *out = (int32_t)days_from_civil(y, (unsigned)m, (unsigned)d); return 1;
}
/* float -> int: Rust `as` (saturating, NaN -> 0)  cast.rs:54-95 */
API int64_t orc_f64_to_int(double v, int bits) {
    if (isnan(v)) return 0;
    double lo = (bits == 64) ? -9223372036854775808.0 : -(double)((int64_t)1 << (bits - 1));
    double hi = (bits == 64) ? 9223372036854775808.0 : (double)((int64_t)1 << (bits - 1));
    if (v <= lo) return (bits == 64) ? INT64_MIN : -((int64_t)1 << (bits - 1));
    if (v >= hi) return (bits == 64) ? INT64_MAX : ((int64_t)1 << (bits - 1)) - 1;
    return (int64_t)v;
}

/* ------------------------------------------------------------------ */
/* CPU baseline legs (timed by bench.py next to the GPU numbers)        */
/* hash aggregate SUM/COUNT by int64 key, open addressing in the spirit */
/* of datafusion-ext-plans/src/agg/agg_hash_map.rs:77-136 (linear probe,*/
/* load <= 0.5) with per-thread tables merged at the end                */
/* (agg_table.rs partial -> merge).  Returns number of groups.          */
/* ------------------------------------------------------------------ */
typedef struct { int64_t* keys; int64_t* sums; int64_t* cnts; uint8_t* used; int64_t cap, n; int has_null; int64_t null_sum, null_cnt; } tbl_t;
static inline uint64_t mixh(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
static void tbl_init(tbl_t* t, int64_t cap) {
    t->cap = cap; t->n = 0; t->has_null = 0; t->null_sum = t->null_cnt = 0;
    t->keys = malloc(cap * 8); t->sums = calloc(cap, 8); t->cnts = calloc(cap, 8); t->used = calloc(cap, 1);
}
static void tbl_free(tbl_t* t) { free(t->keys); free(t->sums); free(t->cnts); free(t->used); }
static void tbl_add(tbl_t* t, int64_t k, int64_t s, int64_t c);
static void tbl_grow(tbl_t* t) {
    tbl_t n; tbl_init(&n, t->cap * 2);
    for (int64_t i = 0; i < t->cap; i++) if (t->used[i]) tbl_add(&n, t->keys[i], t->sums[i], t->cnts[i]);
    n.has_null = t->has_null; n.null_sum = t->null_sum; n.null_cnt = t->null_cnt;
    tbl_free(t); *t = n;
}
static void tbl_add(tbl_t* t, int64_t k, int64_t s, int64_t c) {
    if (t->n * 2 >= t->cap) tbl_grow(t);
    uint64_t m = (uint64_t)t->cap - 1, i = mixh((uint64_t)k) & m;
    while (t->used[i] && t->keys[i] != k) i = (i + 1) & m;
    if (!t->used[i]) { t->used[i] = 1; t->keys[i] = k; t->n++; }
    t->sums[i] += s; t->cnts[i] += c;
}
/* value validity may be NULL; a NULL value contributes nothing to SUM/COUNT (sum.rs:98-123, count.rs:89-125).
 * pred (optional, byte per row) = filter mask applied before aggregation. out arrays sized >= groups. */
#ifdef _OPENMP
#include <omp.h>
#endif
API int64_t orc_agg_sum_count_i64(const int64_t* keys, const uint8_t* key_valid, const int64_t* vals, const uint8_t* val_valid,
                                  const uint8_t* pred, int64_t n, int64_t* out_keys, uint8_t* out_key_valid,
                                  int64_t* out_sums, uint8_t* out_sum_valid, int64_t* out_cnts, int64_t out_cap) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    tbl_t* parts = malloc(sizeof(tbl_t) * nt);
    for (int t = 0; t < nt; t++) tbl_init(&parts[t], 1024);
    #pragma omp parallel num_threads(nt)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        tbl_t* t = &parts[tid];
        int64_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
        for (int64_t i = lo; i < hi; i++) {
            if (pred && !pred[i]) continue;
            int vv = !val_valid || bit_get(val_valid, i);
            if (key_valid && !bit_get(key_valid, i)) { t->has_null = 1; if (vv) { t->null_sum += vals[i]; t->null_cnt++; } continue; }
            tbl_add(t, keys[i], vv ? vals[i] : 0, vv ? 1 : 0);
        }
    }
    tbl_t fin; tbl_init(&fin, 1024);
    for (int t = 0; t < nt; t++) {
        for (int64_t i = 0; i < parts[t].cap; i++) if (parts[t].used[i]) tbl_add(&fin, parts[t].keys[i], parts[t].sums[i], parts[t].cnts[i]);
        if (parts[t].has_null) { fin.has_null = 1; fin.null_sum += parts[t].null_sum; fin.null_cnt += parts[t].null_cnt; }
        tbl_free(&parts[t]);
    }
    free(parts);
    int64_t g = 0;
    for (int64_t i = 0; i < fin.cap && g < out_cap; i++) if (fin.used[i]) {
        out_keys[g] = fin.keys[i]; out_key_valid[g] = 1; out_sums[g] = fin.sums[i]; out_sum_valid[g] = fin.cnts[i] > 0; out_cnts[g] = fin.cnts[i]; g++;
    }
    if (fin.has_null && g < out_cap) { out_keys[g] = 0; out_key_valid[g] = 0; out_sums[g] = fin.null_sum; out_sum_valid[g] = fin.null_cnt > 0; out_cnts[g] = fin.null_cnt; g++; }
    tbl_free(&fin);
    return g;
}

/* counting-sort partitioner: datafusion-ext-plans/src/shuffle/buffered_data.rs:285-353 +
 * datafusion-ext-commons/src/algorithm/rdx_sort.rs:24-74.  Emits row order grouped by
 * partition (order inside a partition is unspecified in the reference; this port is stable)
 * and offsets[N+1]. */
API void orc_partition_rows(const int32_t* part_ids, int64_t n, int32_t num_parts, int32_t* out_rows, int64_t* out_offsets) {
    int64_t* cnt = calloc((size_t)num_parts + 1, 8);
    for (int64_t i = 0; i < n; i++) cnt[part_ids[i] + 1]++;
    for (int32_t p = 0; p < num_parts; p++) cnt[p + 1] += cnt[p];
    memcpy(out_offsets, cnt, ((size_t)num_parts + 1) * 8);
    for (int64_t i = 0; i < n; i++) out_rows[cnt[part_ids[i]]++] = (int32_t)i;
    free(cnt);
}
