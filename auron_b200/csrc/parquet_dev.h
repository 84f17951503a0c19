// parquet_dev.h -- descriptors shared by the Parquet host walker (scan_parquet.cc) and the decode kernels.
#pragma once
#include "common.h"

namespace auron {

struct PqPage {
    const uint8_t* def_ptr;   // definition-level section (hybrid RLE, bit width 1), nullptr if none
    const uint8_t* val_ptr;   // value section (dictionary pages: bit-width byte + hybrid indices ; PLAIN: values)
    int32_t def_len, val_len;
    int32_t num_values;       // rows of the page (flat columns)
    int32_t row_start;        // first output row of the page within the batch
    int32_t encoding;         // parquet Encoding of the values
    int32_t dict_id;          // index into PqColumnArgs::dicts
    int32_t all_null;
    int32_t plain_value_base; // strings: position of this page's first PLAIN value in the chunk value table
};
struct PqDict {
    const uint8_t* data;      // PLAIN-encoded fixed-width dictionary values
    int32_t num_values;
    int32_t value_base;       // strings: position of entry 0 in the value table
};
enum { PQ_MODE_VALUES = 0, PQ_MODE_INDEX = 1 };
struct PqColumnArgs {
    const PqPage* pages;
    const PqDict* dicts;
    int32_t n_pages;
    int32_t phys_type, phys_width, type_length;
    int32_t out_type, out_width;
    int32_t max_def;
    int32_t mode;
    void* out;
    uint32_t* out_valid;
    int32_t* out_idx;
};
struct PqByteSection {
    const uint8_t* ptr;
    int64_t len;
    int32_t num_values;
    int32_t value_base;
};

void pq_decode_pages(Ctx& ctx, const PqColumnArgs& a, const std::vector<PqPage>& host_pages);
ColumnPtr pq_build_value_table(Ctx& ctx, const std::vector<PqByteSection>& secs, int64_t total_values, const DType& type);

}  // namespace auron
