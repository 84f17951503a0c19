// shuffle_writer.cc -- ShuffleWriterExec (rows S2, S4-S6 of SURVEY.md section 8a).  Mirrors
// ShuffleWriterExec::execute (datafusion-ext-plans/src/shuffle_writer_exec.rs:106-168), the repartitioners
// (shuffle/sort_repartitioner.rs:123-254, single_repartitioner.rs:64-97), BufferedData::write
// (shuffle/buffered_data.rs:123-158) and IpcCompressionWriter (datafusion-ext-commons/src/io/ipc_compression.rs:35-113).
//
// Per device chunk: partition ids on device (k_hash.cu), stable counting sort of row ids (k_sort.cu), one gather
// into a partition-contiguous batch, byte-plane serialisation of all partitions in one pass (k_serde.cu), one
// D2H, then LZ4-frame / ZSTD block compression per partition on the host cores (system liblz4 / libzstd via
// dlopen; the reference uses lz4_flex frame / zstd, ipc_compression.rs:178-197).  Output files:
//   data  = for each partition, its blocks from every chunk, concatenated     (sort_repartitioner.rs:217-249)
//   block = u32_le compressed_len | codec stream                              (ipc_compression.rs:84-103)
//   index = (N+1) little-endian i64 offsets, first = 0                        (sort_repartitioner.rs:166-193)
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <thread>

#include "exchange.h"
#include "operators.h"
#include "pb.h"

namespace auron {

namespace {
typedef size_t (*lz4f_bound_fn)(size_t, const void*);
typedef size_t (*lz4f_compress_fn)(void*, size_t, const void*, size_t, const void*);
typedef unsigned (*lz4f_iserror_fn)(size_t);
typedef size_t (*zstd_bound_fn)(size_t);
typedef size_t (*zstd_compress_fn)(void*, size_t, const void*, size_t, int);
typedef unsigned (*zstd_iserror_fn)(size_t);

struct Codecs {
    lz4f_bound_fn lz4_bound = nullptr;
    lz4f_compress_fn lz4_compress = nullptr;
    lz4f_iserror_fn lz4_iserr = nullptr;
    zstd_bound_fn z_bound = nullptr;
    zstd_compress_fn z_compress = nullptr;
    zstd_iserror_fn z_iserr = nullptr;
    Codecs() {
        if (void* h = dlopen("liblz4.so.1", RTLD_NOW)) {
            lz4_bound = (lz4f_bound_fn)dlsym(h, "LZ4F_compressFrameBound");
            lz4_compress = (lz4f_compress_fn)dlsym(h, "LZ4F_compressFrame");
            lz4_iserr = (lz4f_iserror_fn)dlsym(h, "LZ4F_isError");
        }
        if (void* h = dlopen("libzstd.so.1", RTLD_NOW)) {
            z_bound = (zstd_bound_fn)dlsym(h, "ZSTD_compressBound");
            z_compress = (zstd_compress_fn)dlsym(h, "ZSTD_compress");
            z_iserr = (zstd_iserror_fn)dlsym(h, "ZSTD_isError");
        }
    }
};
const Codecs& codecs() {
    static Codecs c;
    return c;
}

// one block: u32_le compressed_len | codec stream
void compress_block(bool zstd, const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    const Codecs& c = codecs();
    size_t bound, written;
    if (zstd) {
        AURON_CHECK(c.z_compress, "libzstd.so.1 not available");
        bound = c.z_bound(n);
        out.resize(4 + bound);
        written = c.z_compress(out.data() + 4, bound, in, n, 1);   // spark.io.compression.zstd.level default 1 (ipc_compression.rs:186-189)
        AURON_CHECK(!c.z_iserr(written), "zstd compression failed");
    } else {
        AURON_CHECK(c.lz4_compress, "liblz4.so.1 not available");
        bound = c.lz4_bound(n, nullptr);
        out.resize(4 + bound);
        written = c.lz4_compress(out.data() + 4, bound, in, n, nullptr);
        AURON_CHECK(!c.lz4_iserr(written), "lz4 frame compression failed");
    }
    uint32_t len = (uint32_t)written;
    memcpy(out.data(), &len, 4);
    out.resize(4 + written);
}
}  // namespace

struct ShuffleWriterExec : Operator {
    int kind = 1;   // 1 single, 2 hash, 3 round robin
    std::vector<ExprPtr> hash_exprs;
    int64_t num_parts = 1;
    std::string data_file, index_file;
    bool done = false;
    bool zstd = false;
    // [chunk][partition] compressed blocks
    std::vector<std::vector<std::vector<uint8_t>>> chunks;
    int64_t rows_so_far = 0;

    void write_chunk(Task& t, const BatchPtr& in) {
        Ctx& ctx = t.ctx;
        int64_t n = in->num_rows;
        std::vector<int64_t> row_off(num_parts + 1, 0);
        BatchPtr sorted = in;
        if (num_parts > 1) {
            Buf pids;
            if (kind == 2) {
                std::vector<ColumnPtr> keys;
                for (auto& e : hash_exprs) keys.push_back(eval_to_column(t, e, children[0]->out_schema, *in));
                pids = murmur3_partition_ids(ctx, keys, n, (int32_t)num_parts, 42);
            } else {
                fail("round-robin / range repartitioning on device is not built yet (hash and single are)");
            }
            Buf rows, offs;
            partition_rows(ctx, P<int32_t>(pids), n, (int32_t)num_parts, &rows, &offs);
            to_host(ctx, row_off.data(), offs->ptr, (size_t)(num_parts + 1) * 8);
            sorted = take_batch(ctx, *in, P<int32_t>(rows), n, false);
        } else {
            row_off[1] = n;
        }
        SerializedParts ser = serialize_partitions(ctx, *sorted, row_off);
        int64_t total = ser.part_offsets.back();
        std::vector<uint8_t> host((size_t)total);
        to_host(ctx, host.data(), ser.bytes->ptr, (size_t)total);
        metrics.add("data_size", total);
        std::vector<std::vector<uint8_t>> blocks((size_t)num_parts);
        // host-side block compression, one partition per work item
        unsigned nthreads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        std::atomic<int64_t> next{0};
        std::string err;
        std::mutex emu;
        auto work = [&]() {
            for (;;) {
                int64_t p = next.fetch_add(1);
                if (p >= num_parts) return;
                int64_t b = ser.part_offsets[p], e = ser.part_offsets[p + 1];
                if (e == b) continue;
                try {
                    compress_block(zstd, host.data() + b, (size_t)(e - b), blocks[(size_t)p]);
                } catch (const std::exception& ex) {
                    std::lock_guard<std::mutex> l(emu);
                    err = ex.what();
                }
            }
        };
        std::vector<std::thread> th;
        for (unsigned i = 1; i < nthreads; i++) th.emplace_back(work);
        work();
        for (auto& x : th) x.join();
        if (!err.empty()) fail(err);
        chunks.push_back(std::move(blocks));
        rows_so_far += n;
    }

    void write_files() {
        FILE* df = fopen(data_file.c_str(), "wb");
        AURON_CHECK(df, "cannot create shuffle data file " + data_file);
        std::vector<int64_t> offsets((size_t)num_parts + 1, 0);
        int64_t pos = 0;
        for (int64_t p = 0; p < num_parts; p++) {
            offsets[(size_t)p] = pos;
            for (auto& ch : chunks) {
                const auto& blk = ch[(size_t)p];
                if (blk.empty()) continue;
                AURON_CHECK(fwrite(blk.data(), 1, blk.size(), df) == blk.size(), "short write on " + data_file);
                pos += (int64_t)blk.size();
            }
        }
        offsets[(size_t)num_parts] = pos;
        fclose(df);
        FILE* xf = fopen(index_file.c_str(), "wb");
        AURON_CHECK(xf, "cannot create shuffle index file " + index_file);
        AURON_CHECK(fwrite(offsets.data(), 8, offsets.size(), xf) == offsets.size(), "short write on " + index_file);
        fclose(xf);
        metrics.add("output_rows", rows_so_far);
    }

    // in-box repartition: output_data_file = "nccl://<name>" turns the writer into an exchange whose output stream is
    // the set of rows this rank owns after the all-to-all (so the next stage can be chained in the same task plan)
    BatchPtr exchange(Task& t) {
        std::vector<BatchPtr> all;
        while (BatchPtr b = children[0]->next(t)) {
            AURON_CHECK(t.is_running(), "task killed");
            if (b->num_rows) all.push_back(b);
        }
        BatchPtr in;
        if (all.empty()) {
            in = std::make_shared<Batch>();
            for (auto& f : out_schema.fields) in->cols.push_back(make_column(t.ctx, f.type, 0, false));
        } else in = concat_batches(t.ctx, all);
        all.clear();
        int64_t n = in->num_rows;
        AURON_CHECK(kind == 2, "the NCCL exchange implements hash repartitioning");
        std::vector<ColumnPtr> keys;
        for (auto& e : hash_exprs) keys.push_back(eval_to_column(t, e, children[0]->out_schema, *in));
        Buf pids = murmur3_partition_ids(t.ctx, keys, n, (int32_t)num_parts, 42);
        Buf rows, offs;
        partition_rows(t.ctx, P<int32_t>(pids), n, (int32_t)num_parts, &rows, &offs);
        std::vector<int64_t> row_off(num_parts + 1, 0);
        to_host(t.ctx, row_off.data(), offs->ptr, (size_t)(num_parts + 1) * 8);
        BatchPtr sorted = take_batch(t.ctx, *in, P<int32_t>(rows), n, false);
        int64_t sent = 0;
        BatchPtr out;
        {
            OpTimer timer(metrics, "exchange_ns");
            out = nccl_exchange(t.ctx, *sorted, row_off, num_parts, &sent);
        }
        metrics.add("data_size", sent);
        metrics.add("output_rows", out->num_rows);
        return out;
    }

    BatchPtr next(Task& t) override {
        if (done) return nullptr;
        done = true;
        if (data_file.rfind("nccl://", 0) == 0) return exchange(t);
        while (BatchPtr b = children[0]->next(t)) {
            AURON_CHECK(t.is_running(), "task killed");
            if (b->num_rows == 0) continue;
            write_chunk(t, b);
        }
        write_files();
        return nullptr;   // the output stream of a shuffle writer is empty (shuffle/mod.rs:61-108)
    }
};

OperatorPtr make_shuffle_writer(Task& t, OperatorPtr input, const uint8_t* node, size_t n) {
    auto op = std::make_unique<ShuffleWriterExec>();
    op->name = "ShuffleWriterExec";
    op->out_schema = input->out_schema;
    PbReader r(node, n);
    uint32_t f, w;
    while (r.next(&f, &w)) {
        if (f == 2 && w == 2) {   // PhysicalRepartition oneof
            const uint8_t* pb;
            size_t pn;
            r.bytes_view(&pb, &pn);
            PbReader p(pb, pn);
            uint32_t pf, pw;
            while (p.next(&pf, &pw)) {
                if (pw != 2) {
                    p.skip(pw);
                    continue;
                }
                const uint8_t* sb;
                size_t sn;
                p.bytes_view(&sb, &sn);
                PbReader s(sb, sn);
                uint32_t sf, sw;
                op->kind = (int)pf;
                AURON_CHECK(pf >= 1 && pf <= 3, "range repartitioning is not built on device yet");
                while (s.next(&sf, &sw)) {
                    if (pf == 2 && sf == 1 && sw == 2) {
                        const uint8_t* eb;
                        size_t en;
                        s.bytes_view(&eb, &en);
                        op->hash_exprs.push_back(decode_expr(eb, en));
                    } else if (((pf == 2 && sf == 2) || (pf != 2 && sf == 1)) && sw == 0) op->num_parts = (int64_t)s.varint();
                    else s.skip(sw);
                }
            }
        } else if (f == 3 && w == 2) op->data_file = r.bytes();
        else if (f == 4 && w == 2) op->index_file = r.bytes();
        else r.skip(w);
    }
    AURON_CHECK(op->num_parts >= 1, "shuffle writer needs at least one output partition");
    if (op->kind == 1) op->num_parts = 1;   // SingleShuffleRepartitioner (single_repartitioner.rs:64-97)
    if (const char* c = getenv("AURON_IO_COMPRESSION_CODEC")) op->zstd = std::string(c) == "zstd";   // spark.io.compression.codec (lz4 | zstd)
    op->children.push_back(std::move(input));
    (void)t;
    return op;
}

}  // namespace auron
