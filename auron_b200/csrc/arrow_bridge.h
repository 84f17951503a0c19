// arrow_bridge.h -- Arrow C Data Interface (the published ABI structs) and the HBM import/export.
// The reference hands batches across its boundary as FFI_ArrowArray / FFI_ArrowSchema
// (auron/src/rt.rs:167-170,258-262 ; datafusion-ext-plans/src/ffi_reader_exec.rs:182-251).
#pragma once
#include "common.h"

extern "C" {
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
    const char* format;
    const char* name;
    const char* metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema** children;
    struct ArrowSchema* dictionary;
    void (*release)(struct ArrowSchema*);
    void* private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void** buffers;
    struct ArrowArray** children;
    struct ArrowArray* dictionary;
    void (*release)(struct ArrowArray*);
    void* private_data;
};
#endif
}

namespace auron {

DType dtype_from_format(const char* format);
std::string format_of(const DType& t);
Schema schema_from_arrow(const ArrowSchema* s);                 // struct schema -> fields
void schema_to_arrow(const Schema& s, ArrowSchema* out);        // caller releases

// H2D: copies the buffers of a struct array (one child per column) into HBM.  Does NOT release `arr`.
BatchPtr import_batch(Ctx& ctx, const ArrowArray* arr, const Schema& schema);
// rows [lo, lo + len) of a host struct array (a spilled sorted run read back range by range)
BatchPtr import_batch_slice(Ctx& ctx, const ArrowArray* arr, const Schema& schema, int64_t lo, int64_t len);
// D2H: materialises host Arrow buffers (malloc'd, freed by the release callback).
// Results of at least `pinned_from` bytes land in a block of the pinned pool (blocks are at least 64 MB), smaller ones in plain memory.
void export_batch(Ctx& ctx, const Batch& b, const Schema& schema, ArrowArray* out, size_t pinned_from = 1u << 20);

}  // namespace auron
