// parquet_dev.h -- descriptors shared by the Parquet host walker (scan_parquet.cc) and the decode kernels.
#pragma once
#include "common.h"

namespace auron {

struct PqPage {
    const uint8_t* def_ptr;   // definition-level section (hybrid RLE, bit width 1), nullptr if none
    const uint8_t* val_ptr;   // value section (dictionary pages: bit-width byte + hybrid indices ; PLAIN: values)
    int32_t def_len, val_len;   // def_len == -1: v1 page decompressed on device, sections not split yet (pq_fix_v1_pages)
    int32_t num_values;       // rows of the page (flat columns)
    int32_t row_start;        // first output row of the page within the batch
    int32_t encoding;         // parquet Encoding of the values
    int32_t dict_id;          // index into PqColumnArgs::dicts
    int32_t all_null;
    int32_t plain_value_base; // strings: position of this page's first PLAIN value in the chunk value table
    int32_t job;              // v1 page decompressed on the device: index of its PqDecompJob / PqDecompResult, else -1
    int32_t pad;
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

// page decompression on device (k_snappy.cu): one warp per job
struct PqDecompJob {
    const uint8_t* src;
    uint8_t* dst;
    int32_t src_len, dst_len;
    int32_t kind;       // 0 = stored bytes, 1 = Snappy raw block
    int32_t v1_levels;  // 1 = body of a nullable v1 data page ([u32 length][levels][values]): a final literal that holds the
                        //     whole value section is NOT copied, the page then reads its values from the compressed buffer
};
struct PqDecompResult {
    const uint8_t* tail_src;   // where the value section lives inside the compressed buffer
    int32_t tail_start;        // its offset in the uncompressed body, -1 = everything was copied
    int32_t pad;
};
struct PqDecompOut {
    Buf status;    // int32: 0 = ok, else 1 + index of the first failing job
    Buf results;   // PqDecompResult per job
};
PqDecompOut pq_decompress(Ctx& ctx, const std::vector<PqDecompJob>& jobs);
// pages with def_len == -1: level / value sections from the body's length word (+ in-place value sections, see above)
void pq_fix_v1_pages(Ctx& ctx, PqPage* pages, int n, const PqDecompResult* results);

// scout + decode of one column; or in three steps, so that one scout launch serves every column of a batch
void pq_decode_pages(Ctx& ctx, const PqColumnArgs& a, const std::vector<PqPage>& host_pages);
struct PqPrepared {
    PqColumnArgs a;
    int n_tiles = 0;
    Buf tile_base, tiles, tile_valid;
};
PqPrepared pq_prepare(Ctx& ctx, const PqColumnArgs& a, const std::vector<PqPage>& host_pages);
void pq_scout_many(Ctx& ctx, const std::vector<PqPrepared*>& cols);
void pq_decode_prepared(Ctx& ctx, const PqPrepared& pr);
ColumnPtr pq_build_value_table(Ctx& ctx, const std::vector<PqByteSection>& secs, int64_t total_values, const DType& type);

}  // namespace auron
