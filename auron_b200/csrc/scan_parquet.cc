// scan_parquet.cc -- ParquetScanExec: host side of the Parquet scan (row P1).  Mirrors ParquetExec::execute
// (datafusion-ext-plans/src/parquet_exec.rs:151-204), the IO adaptor (:294-467: ranged reads through
// FSDataInputWrapper.readFully) and AuronSchemaAdapter (scan/mod.rs:56-160: case-insensitive column match,
// missing columns -> NULL, INT32/INT64 decimals widened by value copy).
//
// Host work: footer + page headers (Thrift), row-group selection by file range, optional host
// decompression of SNAPPY / ZSTD / LZ4_RAW pages.  Everything per value happens in k_parquet.cu.  When the
// file bytes are already resident in HBM (auron_b200_put_device_file) page payloads are decoded in place.
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>

#include "../../include/auron_b200.h"
#include "operators.h"
#include "parquet_dev.h"
#include "parquet_meta.h"
#include "pb.h"

namespace auron {

// ------------------------------------------------------------------------------------------ device-resident files
struct DeviceFile {
    std::vector<uint8_t> host;   // page headers / footer are parsed from the host copy
    std::shared_ptr<Ctx> ctx;    // declared before `dev`: the buffer is freed on this stream, so it must die first
    Buf dev;
};
static std::mutex g_file_mu;
static std::map<std::string, std::shared_ptr<DeviceFile>> g_dev_files;

void put_device_file(const std::string& path, const uint8_t* bytes, size_t len, int device) {
    auto f = std::make_shared<DeviceFile>();
    f->ctx = std::make_shared<Ctx>(device);
    f->host.assign(bytes, bytes + len);
    f->dev = to_device(*f->ctx, bytes, len);
    f->ctx->sync();
    std::lock_guard<std::mutex> l(g_file_mu);
    g_dev_files[path] = f;
}
void drop_device_file(const std::string& path) {
    std::lock_guard<std::mutex> l(g_file_mu);
    g_dev_files.erase(path);
}
static std::shared_ptr<DeviceFile> find_device_file(const std::string& path) {
    std::lock_guard<std::mutex> l(g_file_mu);
    auto it = g_dev_files.find(path);
    return it == g_dev_files.end() ? nullptr : it->second;
}

// ------------------------------------------------------------------------------------------ codecs (system libs, no headers in the image)
typedef size_t (*zstd_decompress_fn)(void*, size_t, const void*, size_t);
typedef unsigned (*zstd_iserror_fn)(size_t);
typedef int (*lz4_decompress_fn)(const char*, char*, int, int);
static void host_decompress(int codec, const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    if (codec == pq::CODEC_SNAPPY) {
        pq::snappy_decompress(in, in_len, out, out_len);
    } else if (codec == pq::CODEC_ZSTD) {
        static void* h = dlopen("libzstd.so.1", RTLD_NOW);
        AURON_CHECK(h, "libzstd.so.1 not available");
        static auto dec = (zstd_decompress_fn)dlsym(h, "ZSTD_decompress");
        static auto iserr = (zstd_iserror_fn)dlsym(h, "ZSTD_isError");
        size_t r = dec(out, out_len, in, in_len);
        AURON_CHECK(!iserr(r) && r == out_len, "zstd page decompression failed");
    } else if (codec == pq::CODEC_LZ4_RAW) {
        static void* h = dlopen("liblz4.so.1", RTLD_NOW);
        AURON_CHECK(h, "liblz4.so.1 not available");
        static auto dec = (lz4_decompress_fn)dlsym(h, "LZ4_decompress_safe");
        int r = dec((const char*)in, (char*)out, (int)in_len, (int)out_len);
        AURON_CHECK(r == (int)out_len, "lz4 page decompression failed");
    } else {
        fail("parquet codec " + std::to_string(codec) + " is not supported (UNCOMPRESSED, SNAPPY, ZSTD, LZ4_RAW are)");
    }
}

// ------------------------------------------------------------------------------------------ operator
struct PqFileSpec {
    std::string path;
    int64_t size = 0, range_start = -1, range_end = -1;
};
struct LeafColumn {
    int leaf_index;
    pq::SchemaElement el;
};

struct ParquetScanExec : Operator {
    std::vector<PqFileSpec> files;
    Schema table_schema;
    std::vector<int> projection;
    std::string fs_id;
    size_t file_pos = 0;
    // current file state
    bool file_open = false;
    std::shared_ptr<DeviceFile> dev_file;
    int fd = -1;
    pq::FileMeta meta;
    std::vector<LeafColumn> leaves;
    std::vector<size_t> row_groups;   // selected row groups of the current file
    size_t rg_pos = 0;
    void* pinned = nullptr;
    size_t pinned_cap = 0;

    ~ParquetScanExec() override {
        if (fd >= 0) close(fd);
        if (pinned) cudaFreeHost(pinned);
    }
    void read_at(Task& t, const PqFileSpec& f, int64_t pos, void* dst, int64_t len) {
        if (dev_file) {
            AURON_CHECK(pos >= 0 && (size_t)(pos + len) <= dev_file->host.size(), "parquet read out of range");
            memcpy(dst, dev_file->host.data() + pos, (size_t)len);
            return;
        }
        if (t.cb && t.cb->read_fully) {   // FSDataInputWrapper.readFully (internal_file_reader.rs:64-68)
            int64_t got = t.cb->read_fully(t.cb->user, fs_id.c_str(), f.path.c_str(), pos, dst, len);
            AURON_CHECK(got == len, "read_fully failed for " + f.path);
            return;
        }
        if (fd < 0) {
            fd = open(f.path.c_str(), O_RDONLY);
            AURON_CHECK(fd >= 0, "cannot open " + f.path);
        }
        int64_t done = 0;
        while (done < len) {
            ssize_t r = pread(fd, (uint8_t*)dst + done, (size_t)(len - done), pos + done);
            AURON_CHECK(r > 0, "short read on " + f.path);
            done += r;
        }
    }
    void* staging(size_t n) {
        if (n > pinned_cap) {
            if (pinned) cudaFreeHost(pinned);
            pinned_cap = std::max<size_t>(n, 64 << 20);
            CUDA_OK(cudaHostAlloc(&pinned, pinned_cap, cudaHostAllocDefault));
        }
        return pinned;
    }
    void open_file(Task& t) {
        const PqFileSpec& f = files[file_pos];
        if (fd >= 0) {
            close(fd);
            fd = -1;
        }
        dev_file = find_device_file(f.path);
        int64_t size = dev_file ? (int64_t)dev_file->host.size() : f.size;
        AURON_CHECK(size >= 12, "not a parquet file: " + f.path);
        uint8_t tail[8];
        read_at(t, f, size - 8, tail, 8);
        AURON_CHECK(memcmp(tail + 4, "PAR1", 4) == 0, "missing PAR1 magic in " + f.path);
        uint32_t flen;
        memcpy(&flen, tail, 4);
        AURON_CHECK((int64_t)flen + 8 <= size, "corrupt parquet footer length");
        std::vector<uint8_t> footer(flen);
        read_at(t, f, size - 8 - flen, footer.data(), flen);
        meta = pq::parse_file_meta(footer.data(), footer.size());
        leaves.clear();
        AURON_CHECK(!meta.schema.empty(), "empty parquet schema");
        int leaf = 0;
        for (size_t i = 1; i < meta.schema.size(); i++) {
            const auto& el = meta.schema[i];
            AURON_CHECK(el.num_children == 0, "nested parquet columns are out of scope (" + el.name + ")");
            AURON_CHECK(el.repetition != 2, "repeated parquet columns are out of scope (" + el.name + ")");
            leaves.push_back({leaf++, el});
        }
        row_groups.clear();
        for (size_t g = 0; g < meta.row_groups.size(); g++) {
            const auto& rg = meta.row_groups[g];
            if (rg.columns.empty()) continue;
            int64_t start = rg.columns[0].start_offset();
            if (f.range_start >= 0 && !(start >= f.range_start && start < f.range_end)) continue;
            row_groups.push_back(g);
        }
        rg_pos = 0;
        file_open = true;
    }
    int find_leaf(const std::string& name) const {
        for (size_t i = 0; i < leaves.size(); i++)
            if (leaves[i].el.name == name) return (int)i;
        for (size_t i = 0; i < leaves.size(); i++) {   // case-insensitive (scan/mod.rs:56-100)
            const std::string& n = leaves[i].el.name;
            if (n.size() != name.size()) continue;
            bool eq = true;
            for (size_t k = 0; k < n.size(); k++) eq = eq && tolower(n[k]) == tolower(name[k]);
            if (eq) return (int)i;
        }
        return -1;
    }

    struct ChunkPages {
        std::vector<PqPage> pages;
        std::vector<PqDict> dicts;
        std::vector<PqByteSection> secs;
        std::vector<Buf> keep;
        int64_t value_table_size = 0;
    };

    // walk the pages of one column chunk, appending page / dictionary descriptors
    void walk_chunk(Task& t, const PqFileSpec& f, const pq::ColumnMeta& cm, const pq::SchemaElement& el, bool is_string, int64_t row_start, ChunkPages& out) {
        int64_t start = cm.start_offset(), len = cm.total_compressed;
        const uint8_t* host;
        const uint8_t* dev;
        if (dev_file) {
            host = dev_file->host.data() + start;
            dev = P<uint8_t>(dev_file->dev) + start;
        } else {
            uint8_t* st = (uint8_t*)staging((size_t)len + 64);
            read_at(t, f, start, st, len);
            Buf d = dalloc(t.ctx, (size_t)len + 64);
            CUDA_OK(cudaMemcpyAsync(d->ptr, st, (size_t)len, cudaMemcpyHostToDevice, t.ctx.stream));
            t.ctx.sync();   // the pinned staging buffer is reused by the next chunk
            out.keep.push_back(d);
            host = st;
            dev = P<uint8_t>(d);
        }
        const int max_def = el.repetition == 1 ? 1 : 0;
        const bool compressed = cm.codec != pq::CODEC_UNCOMPRESSED;
        // compressed chunks: decompress page payloads into one host buffer, upload once
        std::vector<uint8_t> unc;
        struct Fix {
            size_t page;   // index into out.pages, or SIZE_MAX for a dictionary
            size_t dict;
            size_t sec;
            int64_t off;   // offset of the page payload in `unc`
        };
        std::vector<Fix> fixes;
        int64_t pos = 0, values_seen = 0, rows = row_start;
        int cur_dict = -1;
        while (pos < len && values_seen < cm.num_values) {
            pq::PageHeader h = pq::parse_page_header(host + pos, (size_t)(len - pos));
            const uint8_t* payload_h = host + pos + h.header_len;
            const uint8_t* payload_d = dev + pos + h.header_len;
            AURON_CHECK(pos + h.header_len + h.compressed_size <= len, "parquet page overruns its column chunk");
            pos += h.header_len + h.compressed_size;
            if (h.type == pq::PAGE_INDEX) continue;
            int64_t unc_off = -1;
            int32_t lvl_bytes = h.type == pq::PAGE_DATA_V2 ? h.def_bytes + h.rep_bytes : 0;
            if (compressed && !(h.type == pq::PAGE_DATA_V2 && !h.v2_compressed)) {
                unc_off = (int64_t)unc.size();
                unc.resize(unc.size() + (size_t)h.uncompressed_size + 8);
                if (lvl_bytes) memcpy(unc.data() + unc_off, payload_h, (size_t)lvl_bytes);   // v2 levels are never compressed
                host_decompress(cm.codec, payload_h + lvl_bytes, (size_t)(h.compressed_size - lvl_bytes), unc.data() + unc_off + lvl_bytes,
                                (size_t)(h.uncompressed_size - lvl_bytes));
                payload_h = nullptr;   // re-pointed after the upload
            }
            auto hp = [&](int64_t o) -> const uint8_t* { return unc_off >= 0 ? unc.data() + unc_off + o : payload_h + o; };
            if (h.type == pq::PAGE_DICTIONARY) {
                AURON_CHECK(h.encoding == pq::ENC_PLAIN || h.encoding == pq::ENC_PLAIN_DICTIONARY, "unsupported dictionary page encoding");
                PqDict d{payload_d, h.num_values, (int32_t)out.value_table_size};
                cur_dict = (int)out.dicts.size();
                out.dicts.push_back(d);
                size_t sec_idx = SIZE_MAX;
                if (is_string) {
                    sec_idx = out.secs.size();
                    out.secs.push_back({payload_d, h.uncompressed_size, h.num_values, (int32_t)out.value_table_size});
                    out.value_table_size += h.num_values;
                }
                if (unc_off >= 0) fixes.push_back({SIZE_MAX, (size_t)cur_dict, sec_idx, unc_off});
                continue;
            }
            AURON_CHECK(h.type == pq::PAGE_DATA || h.type == pq::PAGE_DATA_V2, "unknown parquet page type");
            PqPage pg;
            memset(&pg, 0, sizeof(pg));
            pg.num_values = h.num_values;
            pg.row_start = (int32_t)rows;
            pg.encoding = h.encoding;
            pg.dict_id = cur_dict;
            AURON_CHECK(h.encoding == pq::ENC_PLAIN || ((h.encoding == pq::ENC_RLE_DICTIONARY || h.encoding == pq::ENC_PLAIN_DICTIONARY) && cur_dict >= 0) ||
                            (h.encoding == pq::ENC_RLE && el.type == pq::PT_BOOLEAN),
                        "parquet encoding " + std::to_string(h.encoding) + " is not supported on device (PLAIN / RLE_DICTIONARY are)");
            int64_t o = 0, total = h.uncompressed_size;
            if (h.type == pq::PAGE_DATA) {
                if (max_def > 0) {
                    AURON_CHECK(h.def_encoding == pq::ENC_RLE, "only RLE definition levels are supported");
                    uint32_t dl;
                    memcpy(&dl, hp(0), 4);
                    pg.def_ptr = (const uint8_t*)(intptr_t)4;   // offsets now, pointers after the base is known
                    pg.def_len = (int32_t)dl;
                    o = 4 + dl;
                }
            } else {
                o = h.rep_bytes;
                if (max_def > 0 && h.def_bytes > 0) {
                    pg.def_ptr = (const uint8_t*)(intptr_t)o;
                    pg.def_len = h.def_bytes;
                }
                if (max_def > 0 && h.def_bytes == 0 && h.num_nulls == h.num_values) pg.all_null = 1;
                o += h.def_bytes;
            }
            AURON_CHECK(o <= total, "corrupt parquet page levels");
            int64_t val_off = o;
            pg.val_len = (int32_t)(total - o);
            const uint8_t* base_d = unc_off >= 0 ? nullptr : payload_d;
            if (base_d) {
                pg.def_ptr = pg.def_len ? base_d + (intptr_t)pg.def_ptr : nullptr;
                pg.val_ptr = base_d + val_off;
            } else {
                pg.val_ptr = (const uint8_t*)(intptr_t)val_off;
            }
            size_t sec_idx = SIZE_MAX;
            if (is_string && h.encoding == pq::ENC_PLAIN) {
                // number of non-null values is only known on device; sections carry the page's value count upper bound
                // => PLAIN string pages need their exact non-null count: v2 gives it, v1 requires the def levels.
                int32_t nn = h.type == pq::PAGE_DATA_V2 ? h.num_values - h.num_nulls : count_non_null_v1(hp(0), max_def, h.num_values);
                pg.plain_value_base = (int32_t)out.value_table_size;
                sec_idx = out.secs.size();
                out.secs.push_back({base_d ? pg.val_ptr : nullptr, pg.val_len, nn, (int32_t)out.value_table_size});
                out.value_table_size += nn;
            }
            if (unc_off >= 0) fixes.push_back({out.pages.size(), 0, sec_idx, unc_off});
            out.pages.push_back(pg);
            rows += h.num_values;
            values_seen += h.num_values;
        }
        if (!unc.empty()) {
            Buf d = to_device(t.ctx, unc.data(), unc.size());
            t.ctx.sync();
            out.keep.push_back(d);
            const uint8_t* base = P<uint8_t>(d);
            for (auto& fx : fixes) {
                if (fx.page == SIZE_MAX) {
                    out.dicts[fx.dict].data = base + fx.off;
                    if (fx.sec != SIZE_MAX) out.secs[fx.sec].ptr = base + fx.off;
                } else {
                    PqPage& pg = out.pages[fx.page];
                    if (pg.def_len) pg.def_ptr = base + fx.off + (intptr_t)pg.def_ptr;
                    else pg.def_ptr = nullptr;
                    intptr_t vo = (intptr_t)pg.val_ptr;
                    pg.val_ptr = base + fx.off + vo;
                    if (fx.sec != SIZE_MAX) out.secs[fx.sec].ptr = pg.val_ptr;
                }
            }
        }
    }
    // host-side count of non-null values of a v1 page (needed only for PLAIN string pages)
    static int32_t count_non_null_v1(const uint8_t* payload, int max_def, int32_t num_values) {
        if (max_def == 0) return num_values;
        uint32_t dl;
        memcpy(&dl, payload, 4);
        const uint8_t* p = payload + 4;
        const uint8_t* end = p + dl;
        int32_t seen = 0, nn = 0;
        while (p < end && seen < num_values) {
            uint32_t h = 0;
            int shift = 0;
            while (p < end) {
                uint8_t b = *p++;
                h |= (uint32_t)(b & 0x7f) << shift;
                if (!(b & 0x80)) break;
                shift += 7;
            }
            if (h & 1) {
                int cnt = (int)(h >> 1) * 8;
                for (int i = 0; i < cnt && seen < num_values; i++, seen++) nn += (p[i >> 3] >> (i & 7)) & 1;
                p += h >> 1;
            } else {
                int cnt = (int)(h >> 1);
                int v = *p++ & 1;
                int take = std::min(cnt, num_values - seen);
                nn += v * take;
                seen += take;
            }
        }
        return nn;
    }

    static int phys_width(int phys, int type_length) {
        switch (phys) {
            case pq::PT_INT32: case pq::PT_FLOAT: return 4;
            case pq::PT_INT64: case pq::PT_DOUBLE: return 8;
            case pq::PT_FLBA: return type_length;
            default: return 0;
        }
    }
    static void check_types(const pq::SchemaElement& el, const DType& t) {
        bool ok = false;
        switch (el.type) {
            case pq::PT_BOOLEAN: ok = t.id == T_BOOL; break;
            case pq::PT_INT32: ok = t.id == T_INT8 || t.id == T_INT16 || t.id == T_INT32 || t.id == T_DATE32 || t.id == T_INT64 || t.id == T_DECIMAL128 || t.id == T_FLOAT64; break;
            case pq::PT_INT64: ok = t.id == T_INT64 || t.id == T_TIMESTAMP || t.id == T_DATE64 || t.id == T_DECIMAL128 || t.id == T_INT32; break;
            case pq::PT_FLOAT: ok = t.id == T_FLOAT32 || t.id == T_FLOAT64; break;
            case pq::PT_DOUBLE: ok = t.id == T_FLOAT64; break;
            case pq::PT_BYTE_ARRAY: ok = t.is_varlen(); break;
            case pq::PT_FLBA: ok = t.id == T_DECIMAL128 && el.type_length <= 16; break;
            default: ok = false;
        }
        AURON_CHECK(ok, "cannot read parquet column " + el.name + " (physical type " + std::to_string(el.type) + ") as " + t.str());
    }

    // per projected column: descriptors accumulated over the row groups (possibly of several files) of one batch
    struct ColState {
        int leaf = -1;   // index into `leaves`, -1 = column missing in the file(s)
        pq::SchemaElement el;
        bool is_string = false;
        ChunkPages cp;
    };
    // layout signature of the current file for the projected columns; a batch never mixes different layouts
    std::string signature() const {
        std::string s;
        for (int pj : projection) {
            int li = find_leaf(table_schema.fields[pj].name);
            if (li < 0) s += "-;";
            else s += std::to_string(leaves[li].leaf_index) + ":" + std::to_string(leaves[li].el.type) + ":" + std::to_string(leaves[li].el.repetition) + ":" +
                      std::to_string(leaves[li].el.type_length) + ";";
        }
        return s;
    }

    BatchPtr build_batch(Task& t, std::vector<ColState>& cols, int64_t n_rows) {
        AURON_CHECK(n_rows < (int64_t)INT32_MAX, "parquet batch too large");
        auto out = std::make_shared<Batch>();
        out->num_rows = n_rows;
        for (size_t ci = 0; ci < projection.size(); ci++) {
            const Field& fld = table_schema.fields[projection[ci]];
            ColState& cs = cols[ci];
            ChunkPages& cp = cs.cp;
            int li = cs.leaf;
            if (li < 0) {   // missing column -> NULL (scan/mod.rs:84-100)
                if (fld.type.is_varlen()) {
                    auto c = make_column(t.ctx, fld.type, n_rows, true);
                    c->null_count = n_rows;
                    CUDA_OK(cudaMemsetAsync(c->offsets->ptr, 0, (size_t)(n_rows + 1) * 4, t.ctx.stream));
                    out->cols.push_back(c);
                } else out->cols.push_back(make_null_column(t.ctx, fld.type, n_rows));
                continue;
            }
            const pq::SchemaElement& el = cs.el;
            const bool is_string = cs.is_string;
            const int max_def = el.repetition == 1 ? 1 : 0;
            PqColumnArgs a;
            memset(&a, 0, sizeof(a));
            Buf dpages = to_device(t.ctx, cp.pages.data(), cp.pages.size() * sizeof(PqPage));
            Buf ddicts = to_device(t.ctx, cp.dicts.empty() ? (const void*)"" : (const void*)cp.dicts.data(), cp.dicts.size() * sizeof(PqDict));
            a.pages = P<PqPage>(dpages);
            a.dicts = P<PqDict>(ddicts);
            a.n_pages = (int)cp.pages.size();
            a.phys_type = el.type;
            a.type_length = el.type_length;
            a.phys_width = phys_width(el.type, el.type_length);
            a.out_type = fld.type.id;
            a.out_width = fld.type.width();
            a.max_def = max_def;
            Buf validity;
            if (max_def > 0) validity = dalloc_zero(t.ctx, bitmap_alloc_bytes(n_rows));
            a.out_valid = P<uint32_t>(validity);
            ColumnPtr col;
            if (is_string) {
                ColumnPtr table = pq_build_value_table(t.ctx, cp.secs, cp.value_table_size, fld.type);
                Buf idx = dalloc(t.ctx, (size_t)std::max<int64_t>(n_rows, 1) * 4);
                a.mode = PQ_MODE_INDEX;
                a.out_idx = P<int32_t>(idx);
                a.out_valid = nullptr;
                pq_decode_pages(t.ctx, a, cp.pages);
                col = take(t.ctx, *table, P<int32_t>(idx), n_rows, max_def > 0);
            } else {
                col = std::make_shared<Column>();
                col->type = fld.type;
                col->len = n_rows;
                if (fld.type.id == T_BOOL) col->data = dalloc_zero(t.ctx, bitmap_alloc_bytes(n_rows));
                else col->data = dalloc(t.ctx, (size_t)n_rows * fld.type.width());
                a.out = col->data->ptr;
                a.mode = PQ_MODE_VALUES;
                pq_decode_pages(t.ctx, a, cp.pages);
                if (validity) {
                    col->validity = validity;
                    col->null_count = -1;
                }
            }
            t.ctx.sync();   // descriptor vectors (host) were uploaded asynchronously
            out->cols.push_back(col);
        }
        return out;
    }

    BatchPtr next(Task& t) override {
        OpTimer timer(metrics, "elapsed_ns");
        std::vector<ColState> cols;
        std::string batch_sig;
        int64_t rows = 0;
        bool started = false;
        for (;;) {
            if (file_pos >= files.size()) break;
            if (!file_open) open_file(t);
            if (rg_pos >= row_groups.size()) {
                file_open = false;
                file_pos++;
                continue;
            }
            const auto& rg = meta.row_groups[row_groups[rg_pos]];
            std::string sig = signature();
            if (started && (sig != batch_sig || rows + rg.num_rows > t.ctx.gpu_chunk_rows)) break;
            AURON_CHECK(t.is_running(), "task killed");
            if (!started) {
                started = true;
                batch_sig = sig;
                cols.assign(projection.size(), ColState());
                for (size_t ci = 0; ci < projection.size(); ci++) {
                    const Field& fld = table_schema.fields[projection[ci]];
                    int li = find_leaf(fld.name);
                    cols[ci].leaf = li;
                    if (li < 0) continue;
                    cols[ci].el = leaves[li].el;
                    check_types(cols[ci].el, fld.type);
                    cols[ci].is_string = cols[ci].el.type == pq::PT_BYTE_ARRAY;
                }
            }
            for (size_t ci = 0; ci < projection.size(); ci++) {
                ColState& cs = cols[ci];
                if (cs.leaf < 0) continue;
                int leaf_index = leaves[find_leaf(table_schema.fields[projection[ci]].name)].leaf_index;
                AURON_CHECK((size_t)leaf_index < rg.columns.size(), "row group misses a column chunk");
                walk_chunk(t, files[file_pos], rg.columns[leaf_index], cs.el, cs.is_string, rows, cs.cp);
            }
            rows += rg.num_rows;
            rg_pos++;
        }
        if (!started) return nullptr;
        OpTimer timer2(metrics, "decode_ns");
        BatchPtr b = build_batch(t, cols, rows);
        metrics.add("output_rows", b->num_rows);
        return b;
    }
};

OperatorPtr make_parquet_scan(Task& t, const uint8_t* node, size_t n) {
    auto op = std::make_unique<ParquetScanExec>();
    op->name = "ParquetExec";
    PbReader r(node, n);
    uint32_t f, w;
    while (r.next(&f, &w)) {
        if (f == 1 && w == 2) {   // FileScanExecConf
            const uint8_t* cb;
            size_t cn;
            r.bytes_view(&cb, &cn);
            PbReader c(cb, cn);
            uint32_t cf, cw;
            while (c.next(&cf, &cw)) {
                if (cf == 3 && cw == 2) {   // FileGroup{files=1}
                    const uint8_t* gb;
                    size_t gn;
                    c.bytes_view(&gb, &gn);
                    PbReader g(gb, gn);
                    uint32_t gf, gw;
                    while (g.next(&gf, &gw)) {
                        if (gf == 1 && gw == 2) {   // PartitionedFile{path=1,size=2,range=5{start=1,end=2}}
                            const uint8_t* fb;
                            size_t fn;
                            g.bytes_view(&fb, &fn);
                            PbReader pf(fb, fn);
                            uint32_t pff, pfw;
                            PqFileSpec spec;
                            while (pf.next(&pff, &pfw)) {
                                if (pff == 1 && pfw == 2) spec.path = pf.bytes();
                                else if (pff == 2 && pfw == 0) spec.size = (int64_t)pf.varint();
                                else if (pff == 5 && pfw == 2) {
                                    const uint8_t* rb;
                                    size_t rn;
                                    pf.bytes_view(&rb, &rn);
                                    PbReader rr(rb, rn);
                                    uint32_t rf, rw;
                                    spec.range_start = 0;
                                    spec.range_end = 0;
                                    while (rr.next(&rf, &rw)) {
                                        if (rf == 1 && rw == 0) spec.range_start = (int64_t)rr.varint();
                                        else if (rf == 2 && rw == 0) spec.range_end = (int64_t)rr.varint();
                                        else rr.skip(rw);
                                    }
                                } else pf.skip(pfw);
                            }
                            op->files.push_back(spec);
                        } else g.skip(gw);
                    }
                } else if (cf == 4 && cw == 2) {
                    const uint8_t* sb;
                    size_t sn;
                    c.bytes_view(&sb, &sn);
                    op->table_schema = decode_schema(sb, sn);
                } else if (cf == 6 && cw == 0) op->projection.push_back((int)c.varint());
                else if (cf == 6 && cw == 2) {   // packed
                    const uint8_t* pb;
                    size_t pn;
                    c.bytes_view(&pb, &pn);
                    PbReader p(pb, pn);
                    while (!p.done()) op->projection.push_back((int)p.varint());
                } else if (cf == 9 && cw == 2) {
                    const uint8_t* sb;
                    size_t sn;
                    c.bytes_view(&sb, &sn);
                    AURON_CHECK(decode_schema(sb, sn).fields.empty(), "hive partition columns are not supported by the device scan yet");
                } else c.skip(cw);
            }
        } else if (f == 3 && w == 2) op->fs_id = r.bytes();
        else r.skip(w);   // pruning_predicates: row-group pruning is an optimisation, results are identical without it
    }
    if (op->projection.empty())
        for (size_t i = 0; i < op->table_schema.fields.size(); i++) op->projection.push_back((int)i);
    for (int p : op->projection) {
        AURON_CHECK(p >= 0 && p < (int)op->table_schema.fields.size(), "scan projection out of range");
        op->out_schema.fields.push_back(op->table_schema.fields[p]);
    }
    (void)t;
    return op;
}

}  // namespace auron
