// rowkeys.cuh -- device-side row hash / row equality over a set of key columns (general path of
// hash aggregation and hash join).  Equality follows the reference's grouping / join comparators:
// grouping keys compare arrow-row bytes (NULL == NULL; datafusion-ext-plans/src/agg/agg_ctx.rs:233-245),
// joins re-verify values with EqComparator (datafusion-ext-commons/src/arrow/eq_comparator.rs:42-98) and
// never match NULL keys (handled by the caller).  Floats compare by bits after the reference's upstream
// NormalizeNaNAndZero, i.e. bitwise equality.
#pragma once
#include "device_utils.cuh"
#include "kernels.h"

namespace auron {

__device__ __forceinline__ uint64_t rowkey_hash(const RowKeys& k, int64_t row) {
    uint64_t h = 0x3F6F1B93ull;
    for (int c = 0; c < k.ncols; c++) {
        const KeyColDesc& d = k.c[c];
        if (d.validity && !bit_get(d.validity, row)) {
            h = mix64(h ^ 0x9e3779b97f4a7c15ull);
            continue;
        }
        uint64_t v;
        switch (d.width) {
            case 1: v = ((const uint8_t*)d.data)[row]; break;
            case 2: v = ((const uint16_t*)d.data)[row]; break;
            case 4: v = ((const uint32_t*)d.data)[row]; break;
            case 8: v = ((const uint64_t*)d.data)[row]; break;
            case 16: {
                ulonglong2 t = ((const ulonglong2*)d.data)[row];
                v = t.x ^ mix64(t.y);
                break;
            }
            default:
                if (d.type == T_BOOL) v = bit_get((const uint8_t*)d.data, row);
                else {
                    int32_t b = d.offsets[row], e = d.offsets[row + 1];
                    v = xxhash64_bytes((const uint8_t*)d.data + b, e - b, 0);
                }
        }
        h = mix64(h ^ v) + 0x632be59bd9b4e019ull * (uint64_t)(c + 1);
    }
    return h;
}

// any key column NULL at this row?
__device__ __forceinline__ bool rowkey_has_null(const RowKeys& k, int64_t row) {
    for (int c = 0; c < k.ncols; c++)
        if (k.c[c].validity && !bit_get(k.c[c].validity, row)) return true;
    return false;
}

// equality of row ra in key set a with row rb in key set b (same column types); NULL == NULL
__device__ __forceinline__ bool rowkey_equal(const RowKeys& a, int64_t ra, const RowKeys& b, int64_t rb) {
    for (int c = 0; c < a.ncols; c++) {
        const KeyColDesc& x = a.c[c];
        const KeyColDesc& y = b.c[c];
        bool vx = !x.validity || bit_get(x.validity, ra), vy = !y.validity || bit_get(y.validity, rb);
        if (vx != vy) return false;
        if (!vx) continue;
        switch (x.width) {
            case 1: if (((const uint8_t*)x.data)[ra] != ((const uint8_t*)y.data)[rb]) return false; break;
            case 2: if (((const uint16_t*)x.data)[ra] != ((const uint16_t*)y.data)[rb]) return false; break;
            case 4: if (((const uint32_t*)x.data)[ra] != ((const uint32_t*)y.data)[rb]) return false; break;
            case 8: if (((const uint64_t*)x.data)[ra] != ((const uint64_t*)y.data)[rb]) return false; break;
            case 16: {
                ulonglong2 p = ((const ulonglong2*)x.data)[ra], q = ((const ulonglong2*)y.data)[rb];
                if (p.x != q.x || p.y != q.y) return false;
                break;
            }
            default:
                if (x.type == T_BOOL) {
                    if (bit_get((const uint8_t*)x.data, ra) != bit_get((const uint8_t*)y.data, rb)) return false;
                } else {
                    int32_t xb = x.offsets[ra], xe = x.offsets[ra + 1], yb = y.offsets[rb], ye = y.offsets[rb + 1];
                    if (xe - xb != ye - yb) return false;
                    const uint8_t* p = (const uint8_t*)x.data + xb;
                    const uint8_t* q = (const uint8_t*)y.data + yb;
                    for (int32_t i = 0; i < xe - xb; i++)
                        if (p[i] != q[i]) return false;
                }
        }
    }
    return true;
}

}  // namespace auron
