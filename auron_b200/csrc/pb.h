// pb.h -- minimal proto3 wire-format reader (no protoc in this image) used to decode the subset of
// auron.proto (auron-planner/proto/auron.proto) that reaches the hot path, plus a tiny flatbuffers
// reader for the Arrow IPC stream that carries ScalarValue literals (auron.proto:879-881;
// auron-planner/src/lib.rs:446-456 reads them with StreamReader).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

#include "common.h"

namespace auron {

struct PbReader {
    const uint8_t* p;
    const uint8_t* end;
    PbReader(const uint8_t* b, size_t n) : p(b), end(b + n) {}
    PbReader(const std::string& s) : p((const uint8_t*)s.data()), end((const uint8_t*)s.data() + s.size()) {}
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        while (p < end) {
            uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
            if (shift > 63) break;
        }
        fail("protobuf: malformed varint");
    }
    // reads the next tag; returns false at end
    bool next(uint32_t* field, uint32_t* wire) {
        if (p >= end) return false;
        uint64_t t = varint();
        *field = (uint32_t)(t >> 3);
        *wire = (uint32_t)(t & 7);
        return true;
    }
    std::string bytes() {
        uint64_t n = varint();
        AURON_CHECK((uint64_t)(end - p) >= n, "protobuf: truncated length-delimited field");
        std::string s((const char*)p, (size_t)n);
        p += n;
        return s;
    }
    // view without copying
    void bytes_view(const uint8_t** b, size_t* n) {
        uint64_t len = varint();
        AURON_CHECK((uint64_t)(end - p) >= len, "protobuf: truncated length-delimited field");
        *b = p;
        *n = (size_t)len;
        p += len;
    }
    void skip(uint32_t wire) {
        switch (wire) {
            case 0: varint(); break;
            case 1: AURON_CHECK(end - p >= 8, "protobuf: truncated fixed64"); p += 8; break;
            case 2: {
                uint64_t n = varint();
                AURON_CHECK((uint64_t)(end - p) >= n, "protobuf: truncated field");
                p += n;
                break;
            }
            case 5: AURON_CHECK(end - p >= 4, "protobuf: truncated fixed32"); p += 4; break;
            default: fail("protobuf: unsupported wire type");
        }
    }
};

// ---- flatbuffers (read-only, just what an Arrow IPC Schema / RecordBatch message needs) ----
// The bytes come from outside (a plan literal): every offset is checked against the message before it is followed, a
// malformed buffer is an error, never an out-of-bounds read.
struct FbTable {
    const uint8_t* base = nullptr;   // message start
    const uint8_t* tbl = nullptr;    // table position
    size_t size = 0;                 // message length
    bool ok() const { return tbl != nullptr; }
    template <typename T>
    static T rd(const uint8_t* p) {
        T v;
        memcpy(&v, p, sizeof(T));
        return v;
    }
    // [p, p + n) lies inside the message
    const uint8_t* chk(const uint8_t* p, size_t n) const {
        const uintptr_t b = (uintptr_t)base, q = (uintptr_t)p;
        AURON_CHECK(base && q >= b && q - b <= size && n <= size - (q - b), "Arrow IPC: flatbuffer offset outside the message");
        return p;
    }
    template <typename T>
    T rdc(const uint8_t* p) const {
        return rd<T>(chk(p, sizeof(T)));
    }
    static FbTable root(const uint8_t* msg, size_t len) {
        FbTable t;
        t.base = msg;
        t.size = len;
        AURON_CHECK(len >= 8, "Arrow IPC: message too short");
        t.tbl = msg + rd<uint32_t>(msg);
        t.chk(t.tbl, 4);
        return t;
    }
    // offset of field `id` inside the table, or 0 if absent
    uint16_t field_off(int id) const {
        const int32_t vt_rel = rdc<int32_t>(tbl);
        const uint8_t* vt = (const uint8_t*)((intptr_t)tbl - (intptr_t)vt_rel);
        const uint16_t vt_size = rdc<uint16_t>(vt);
        const uint16_t pos = (uint16_t)(4 + 2 * id);
        if (pos + 2 > vt_size) return 0;
        return rdc<uint16_t>(vt + pos);
    }
    template <typename T>
    T scalar(int id, T def) const {
        uint16_t o = field_off(id);
        return o ? rdc<T>(tbl + o) : def;
    }
    FbTable table(int id) const {
        uint16_t o = field_off(id);
        FbTable t;
        if (!o) return t;
        const uint8_t* p = tbl + o;
        t.base = base;
        t.size = size;
        t.tbl = p + rdc<uint32_t>(p);
        chk(t.tbl, 4);
        return t;
    }
    std::string str(int id) const {
        uint16_t o = field_off(id);
        if (!o) return "";
        const uint8_t* p = tbl + o;
        p += rdc<uint32_t>(p);
        const uint32_t n = rdc<uint32_t>(p);
        return std::string((const char*)chk(p + 4, n), n);
    }
    // vector of `elem`-byte elements: returns pointer to the first element and the count
    const uint8_t* vec(int id, uint32_t* n, size_t elem = 4) const {
        uint16_t o = field_off(id);
        if (!o) {
            *n = 0;
            return nullptr;
        }
        const uint8_t* p = tbl + o;
        p += rdc<uint32_t>(p);
        *n = rdc<uint32_t>(p);
        return chk(p + 4, (size_t)*n * elem);
    }
    FbTable vec_table(const uint8_t* elems, uint32_t i) const {
        FbTable t;
        const uint8_t* p = elems + 4 * (size_t)i;
        t.base = base;
        t.size = size;
        t.tbl = p + rdc<uint32_t>(p);
        chk(t.tbl, 4);
        return t;
    }
};

}  // namespace auron
