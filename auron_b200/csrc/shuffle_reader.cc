// shuffle_reader.cc -- IpcReaderExec: the read side of Auron's compacted shuffle format.
//
// Mirrors datafusion-ext-plans/src/ipc_reader_exec.rs:166-275: the host runtime hands over an iterator of blocks (file
// segment | in-memory buffer; AuronBlockObject.hasFileSegment / hasByteBuffer), every block is a sequence of
//     u32_le compressed_len | codec stream                (IpcCompressionReader, ipc_compression.rs:115-176)
// whose concatenated payload is a sequence of batches in the byte-plane format (read_batch, batch_serde.rs:81-101).
//
// Host side (this file): fetch the blocks into pinned memory, decompress the codec streams on the worker pool (LZ4 frame /
// ZSTD through the system libraries, detected by magic), walk the section layout of every batch (varints, section sizes,
// string byte totals) and ship payload + descriptors to the GPU.  Device side (k_serde.cu): all batches of the chunk become
// one Arrow batch -- planes -> values, validity / bool bits re-packed at the row offset, string lengths -> offsets by a
// scan, string bytes by cooperative copies.  The reference coalesces staged batches up to the batch size
// (ipc_reader_exec.rs:241-262); here a chunk is 256 MB of compressed blocks.
#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>

#include <cstring>

#include "../../include/auron_b200.h"
#include "host_pool.h"
#include "operators.h"
#include "pb.h"

namespace auron {

namespace {
// ---- codecs (system libraries, no headers in this image)
struct ZInBuf {
    const void* src;
    size_t size, pos;
};
struct ZOutBuf {
    void* dst;
    size_t size, pos;
};
struct DecCodecs {
    // LZ4F
    size_t (*lz4_create)(void**, unsigned) = nullptr;
    size_t (*lz4_free)(void*) = nullptr;
    size_t (*lz4_decompress)(void*, void*, size_t*, const void*, size_t*, const void*) = nullptr;
    unsigned (*lz4_iserr)(size_t) = nullptr;
    // ZSTD
    void* (*z_create)() = nullptr;
    size_t (*z_free)(void*) = nullptr;
    size_t (*z_stream)(void*, ZOutBuf*, ZInBuf*) = nullptr;
    unsigned (*z_iserr)(size_t) = nullptr;
    unsigned long long (*z_content_size)(const void*, size_t) = nullptr;
};
const DecCodecs& dec_codecs() {
    static DecCodecs c = [] {
        DecCodecs c;
        if (void* h = dlopen("liblz4.so.1", RTLD_NOW | RTLD_GLOBAL)) {
            c.lz4_create = (decltype(c.lz4_create))dlsym(h, "LZ4F_createDecompressionContext");
            c.lz4_free = (decltype(c.lz4_free))dlsym(h, "LZ4F_freeDecompressionContext");
            c.lz4_decompress = (decltype(c.lz4_decompress))dlsym(h, "LZ4F_decompress");
            c.lz4_iserr = (decltype(c.lz4_iserr))dlsym(h, "LZ4F_isError");
        }
        if (void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL)) {
            c.z_create = (decltype(c.z_create))dlsym(h, "ZSTD_createDStream");
            c.z_free = (decltype(c.z_free))dlsym(h, "ZSTD_freeDStream");
            c.z_stream = (decltype(c.z_stream))dlsym(h, "ZSTD_decompressStream");
            c.z_iserr = (decltype(c.z_iserr))dlsym(h, "ZSTD_isError");
            c.z_content_size = (decltype(c.z_content_size))dlsym(h, "ZSTD_getFrameContentSize");
        }
        return c;
    }();
    return c;
}
// one codec stream -> bytes appended to out
void decompress_stream(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    AURON_CHECK(n >= 4, "shuffle read: truncated codec stream");
    uint32_t magic;
    memcpy(&magic, in, 4);
    const DecCodecs& c = dec_codecs();
    const size_t kStep = 1 << 20;
    if (magic == 0x184D2204u) {
        AURON_CHECK(c.lz4_create && c.lz4_decompress, "liblz4.so.1 not available");
        void* ctx = nullptr;
        AURON_CHECK(!c.lz4_iserr(c.lz4_create(&ctx, 100)), "LZ4F context");
        size_t ip = 0;
        try {
            for (;;) {
                size_t have = out.size();
                out.resize(have + kStep);
                size_t dn = kStep, sn = n - ip;
                size_t r = c.lz4_decompress(ctx, out.data() + have, &dn, in + ip, &sn, nullptr);
                out.resize(have + dn);
                AURON_CHECK(!c.lz4_iserr(r), "shuffle read: corrupt LZ4 frame");
                ip += sn;
                if (r == 0) break;                                   // frame complete
                AURON_CHECK(dn > 0 || sn > 0, "shuffle read: truncated LZ4 frame");
            }
        } catch (...) {
            c.lz4_free(ctx);
            throw;
        }
        c.lz4_free(ctx);
    } else if (magic == 0xFD2FB528u) {
        AURON_CHECK(c.z_create && c.z_stream, "libzstd.so.1 not available");
        void* ds = c.z_create();
        AURON_CHECK(ds, "ZSTD stream");
        ZInBuf ib{in, n, 0};
        try {
            for (;;) {
                size_t have = out.size();
                out.resize(have + kStep);
                ZOutBuf ob{out.data() + have, kStep, 0};
                const size_t before = ib.pos;
                size_t r = c.z_stream(ds, &ob, &ib);
                out.resize(have + ob.pos);
                AURON_CHECK(!c.z_iserr(r), "shuffle read: corrupt ZSTD stream");
                if (r == 0) break;                                   // frame complete
                AURON_CHECK(ob.pos > 0 || ib.pos > before, "shuffle read: truncated ZSTD stream");
            }
        } catch (...) {
            c.z_free(ds);
            throw;
        }
        c.z_free(ds);
    } else {
        fail("shuffle read: unknown codec stream (neither an LZ4 frame nor ZSTD)");
    }
}
// upper bound of a codec stream's decompressed size without decoding it, or -1 when the frame does not say
int64_t stream_bound(const uint8_t* in, size_t n) {
    if (n < 7) return -1;
    uint32_t magic;
    memcpy(&magic, in, 4);
    if (magic == 0x184D2204u) {   // LZ4 frame: header, then { u32 size | data [| u32 checksum] }* until the end mark
        const uint8_t flg = in[4], bd = in[5];
        static const int64_t kMax[8] = {0, 0, 0, 0, 64 << 10, 256 << 10, 1 << 20, 4 << 20};
        const int64_t block_max = kMax[(bd >> 4) & 7];
        if (!block_max) return -1;
        size_t pos = 6;
        if (flg & 0x08) {   // content size present: exact
            if (pos + 8 > n) return -1;
            uint64_t cs;
            memcpy(&cs, in + pos, 8);
            return (int64_t)cs;
        }
        if (flg & 0x01) pos += 4;   // dictionary id
        pos += 1;                   // header checksum
        int64_t blocks = 0;
        for (;;) {
            if (pos + 4 > n) return -1;
            uint32_t w;
            memcpy(&w, in + pos, 4);
            pos += 4;
            if (w == 0) break;
            pos += (w & 0x7fffffffu) + ((flg & 0x10) ? 4 : 0);
            blocks++;
        }
        return blocks * block_max;
    }
    if (magic == 0xFD2FB528u) {
        const DecCodecs& c = dec_codecs();
        if (!c.z_content_size) return -1;
        unsigned long long v = c.z_content_size(in, n);
        return v >= 0xfffffffffffffffeull ? -1 : (int64_t)v;
    }
    return -1;
}
// decode one codec stream straight into dst[0, cap); returns the number of bytes produced
int64_t decompress_stream_into(const uint8_t* in, size_t n, uint8_t* dst, int64_t cap) {
    uint32_t magic;
    memcpy(&magic, in, 4);
    const DecCodecs& c = dec_codecs();
    if (magic == 0x184D2204u) {
        AURON_CHECK(c.lz4_create && c.lz4_decompress, "liblz4.so.1 not available");
        void* ctx = nullptr;
        AURON_CHECK(!c.lz4_iserr(c.lz4_create(&ctx, 100)), "LZ4F context");
        size_t ip = 0;
        int64_t op = 0;
        try {
            for (;;) {
                size_t dn = (size_t)(cap - op), sn = n - ip;
                size_t r = c.lz4_decompress(ctx, dst + op, &dn, in + ip, &sn, nullptr);
                AURON_CHECK(!c.lz4_iserr(r), "shuffle read: corrupt LZ4 frame");
                ip += sn;
                op += (int64_t)dn;
                if (r == 0) break;
                AURON_CHECK(dn > 0 || sn > 0, "shuffle read: truncated LZ4 frame (or a wrong size bound)");
            }
        } catch (...) {
            c.lz4_free(ctx);
            throw;
        }
        c.lz4_free(ctx);
        return op;
    }
    AURON_CHECK(magic == 0xFD2FB528u && c.z_create && c.z_stream, "shuffle read: unknown codec stream");
    void* ds = c.z_create();
    AURON_CHECK(ds, "ZSTD stream");
    ZInBuf ib{in, n, 0};
    ZOutBuf ob{dst, (size_t)cap, 0};
    try {
        for (;;) {
            const size_t b_in = ib.pos, b_out = ob.pos;
            size_t r = c.z_stream(ds, &ob, &ib);
            AURON_CHECK(!c.z_iserr(r), "shuffle read: corrupt ZSTD stream");
            if (r == 0) break;
            AURON_CHECK(ob.pos > b_out || ib.pos > b_in, "shuffle read: truncated ZSTD stream (or a wrong size bound)");
        }
    } catch (...) {
        c.z_free(ds);
        throw;
    }
    c.z_free(ds);
    return (int64_t)ob.pos;
}
uint64_t read_varint(const uint8_t* p, int64_t n, int64_t* pos) {   // io/mod.rs:71-84
    uint64_t v = 0;
    int shift = 0;
    for (;;) {
        AURON_CHECK(*pos < n && shift < 64, "shuffle read: truncated varint");
        uint8_t b = p[(*pos)++];
        v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return v;
        shift += 7;
    }
}
}  // namespace

struct IpcReaderExec : Operator {
    std::string resource_id;
    std::string describe() const override { return "\"resource_id\":" + json_quote(resource_id); }
    bool done = false;
    struct Block {
        std::string path;
        int64_t offset = 0, length = 0;
        const uint8_t* data = nullptr;
        std::vector<uint8_t> copy;   // in-memory blocks are only valid until the next upcall: copied at once
    };

    bool next_block(Task& t, Block* out) {
        AURON_CHECK(t.cb && t.cb->next_shuffle_block, "IpcReaderExec needs the next_shuffle_block callback");
        auron_shuffle_block b;
        memset(&b, 0, sizeof(b));
        int r = t.cb->next_shuffle_block(t.cb->user, resource_id.c_str(), &b);
        AURON_CHECK(r >= 0, "next_shuffle_block failed for resource " + resource_id);
        if (r == 0) return false;
        out->path = b.path ? b.path : "";
        out->offset = b.offset;
        out->length = b.length;
        AURON_CHECK(out->length >= 0 && (b.data || !out->path.empty()), "next_shuffle_block returned an empty block descriptor");
        if (b.data) {
            out->copy.assign(b.data, b.data + out->length);
            out->data = out->copy.data();
        }
        return true;
    }

    struct StreamRef {
        const uint8_t* p;
        size_t n;
    };
    // GPU path: every codec stream is an LZ4 frame with independent blocks -> upload the COMPRESSED bytes, decode the blocks,
    // walk the layout and rebuild the columns on the device; the payload never visits the host.  Returns nullptr when the chunk
    // does not qualify (ZSTD, linked blocks, dictionaries, blocks that are not filled to the block size, batches that continue
    // in the next stream): the host path then handles it.
    BatchPtr try_device_decode(Task& t, const std::vector<StreamRef>& streams, const uint8_t* cbuf, size_t cbuf_bytes) {
        if (streams.empty() || getenv("AURON_HOST_LZ4_DECODE")) return nullptr;
        const int ncols = (int)out_schema.fields.size();
        if (ncols > 64) return nullptr;
        struct Frame {
            size_t first;
            int nblocks;
            int64_t block_max;
        };
        std::vector<Frame> frames;
        std::vector<Lz4DBlock> blocks;
        static const int64_t kMax[8] = {0, 0, 0, 0, 64 << 10, 256 << 10, 1 << 20, 4 << 20};
        for (auto& st : streams) {
            const uint8_t* in = st.p;
            const size_t n = st.n;
            if (n < 11) return nullptr;
            uint32_t magic;
            memcpy(&magic, in, 4);
            if (magic != 0x184D2204u) return nullptr;
            const uint8_t flg = in[4], bd = in[5];
            if ((flg >> 6) != 1 || !(flg & 0x20) || (flg & 0x01)) return nullptr;   // version 1, independent blocks, no dictionary
            const int64_t block_max = kMax[(bd >> 4) & 7];
            if (!block_max) return nullptr;
            size_t pos = 6 + ((flg & 0x08) ? 8 : 0) + 1;
            Frame f{blocks.size(), 0, block_max};
            for (;;) {
                if (pos + 4 > n) return nullptr;
                uint32_t w;
                memcpy(&w, in + pos, 4);
                pos += 4;
                if (w == 0) break;
                const uint32_t len = w & 0x7fffffffu;
                if (pos + len + ((flg & 0x10) ? 4 : 0) > n || (int64_t)len > block_max + 64) return nullptr;
                blocks.push_back(Lz4DBlock{in + pos, nullptr, (int32_t)len, (int32_t)block_max, (w >> 31) ? 1 : 0, 0});
                pos += len + ((flg & 0x10) ? 4 : 0);
                f.nblocks++;
            }
            frames.push_back(f);
        }
        const size_t ns = streams.size();
        std::vector<int64_t> slot(ns + 1, 0);
        for (size_t i = 0; i < ns; i++) slot[i + 1] = slot[i] + (((int64_t)frames[i].nblocks * frames[i].block_max + 63) & ~(int64_t)63);
        if (slot[ns] > (24ll << 30) || blocks.empty()) return nullptr;
        Ctx& ctx = t.ctx;
        OpTimer tdev(metrics, "device_ns");
        Buf dcomp = dalloc(ctx, cbuf_bytes + 64);
        CUDA_OK(cudaMemcpyAsync(dcomp->ptr, cbuf, cbuf_bytes, cudaMemcpyHostToDevice, ctx.stream));
        Buf dpayload = dalloc(ctx, (size_t)slot[ns] + 64);
        for (size_t i = 0; i < ns; i++)
            for (int j = 0; j < frames[i].nblocks; j++) {
                Lz4DBlock& b = blocks[frames[i].first + (size_t)j];
                b.src = P<uint8_t>(dcomp) + (b.src - cbuf);
                b.dst = P<uint8_t>(dpayload) + slot[i] + (int64_t)j * frames[i].block_max;
            }
        Buf dblocks = to_device(ctx, blocks.data(), blocks.size() * sizeof(Lz4DBlock));
        Buf dsizes = dalloc(ctx, blocks.size() * 4);
        lz4_decompress_blocks(ctx, P<Lz4DBlock>(dblocks), (int)blocks.size(), P<int32_t>(dsizes));
        std::vector<int32_t> sizes(blocks.size());
        to_host(ctx, sizes.data(), dsizes->ptr, sizes.size() * 4);
        std::vector<LayoutStream> lstreams(ns);
        int64_t payload_bytes = 0;
        for (size_t i = 0; i < ns; i++) {
            int64_t sz = 0;
            for (int j = 0; j < frames[i].nblocks; j++) {
                const int32_t bs = sizes[frames[i].first + (size_t)j];
                AURON_CHECK(bs >= 0, "shuffle read: corrupt LZ4 block");
                if (j + 1 < frames[i].nblocks && bs != frames[i].block_max) return nullptr;   // a flushed (partial) block in the middle
                sz += bs;
            }
            lstreams[i] = LayoutStream{slot[i], slot[i] + sz};
            payload_bytes += sz;
        }
        // layout: count, then fill
        LayoutSchema sch;
        memset(&sch, 0, sizeof(sch));
        sch.ncols = ncols;
        for (int c = 0; c < ncols; c++) {
            const DType& ty = out_schema.fields[(size_t)c].type;
            sch.kind[c] = ty.id == T_NULL ? 0 : ty.id == T_BOOL ? 1 : ty.is_varlen() ? 3 : 2;
            sch.width[c] = (uint8_t)(sch.kind[c] == 2 ? ty.width() : 0);
        }
        Buf dstreams = to_device(ctx, lstreams.data(), ns * sizeof(LayoutStream));
        Buf dcounts = dalloc(ctx, ns * 4), dflags = dalloc(ctx, ns * 4);
        deserialize_layout(ctx, P<uint8_t>(dpayload), P<LayoutStream>(dstreams), (int)ns, sch, nullptr, nullptr, nullptr, nullptr, P<int32_t>(dcounts), P<int32_t>(dflags));
        std::vector<int32_t> counts(ns), flags(ns);
        to_host(ctx, counts.data(), dcounts->ptr, ns * 4);
        to_host(ctx, flags.data(), dflags->ptr, ns * 4);
        std::vector<int32_t> seg_base(ns + 1, 0);
        for (size_t i = 0; i < ns; i++) {
            if (flags[i]) return nullptr;   // malformed, or a batch that continues in the next stream: the host walk decides
            seg_base[i + 1] = seg_base[i] + counts[i];
        }
        const size_t nbatches = (size_t)seg_base[ns];
        std::vector<DeserSeg> hsegs(nbatches * (size_t)ncols);
        std::vector<int64_t> hbytes(nbatches * (size_t)ncols), hrows(nbatches);
        if (nbatches) {
            Buf dbase = to_device(ctx, seg_base.data(), seg_base.size() * 4);
            Buf dsegs = dalloc_zero(ctx, hsegs.size() * sizeof(DeserSeg)), dbytes = dalloc_zero(ctx, hbytes.size() * 8), drows = dalloc_zero(ctx, nbatches * 8);
            deserialize_layout(ctx, P<uint8_t>(dpayload), P<LayoutStream>(dstreams), (int)ns, sch, P<int32_t>(dbase), P<DeserSeg>(dsegs), P<int64_t>(dbytes), P<int64_t>(drows),
                               P<int32_t>(dcounts), P<int32_t>(dflags));
            to_host(ctx, hsegs.data(), dsegs->ptr, hsegs.size() * sizeof(DeserSeg));
            to_host(ctx, hbytes.data(), dbytes->ptr, hbytes.size() * 8);
            to_host(ctx, hrows.data(), drows->ptr, nbatches * 8);
        }
        std::vector<std::vector<DeserSeg>> segs((size_t)ncols);
        std::vector<std::vector<DeserCopy>> copies((size_t)ncols);
        std::vector<int64_t> col_bytes((size_t)ncols, 0);
        int64_t rows = 0;
        for (size_t bi = 0; bi < nbatches; bi++) {
            for (int c = 0; c < ncols; c++) {
                if (sch.kind[c] == 0) continue;
                DeserSeg sg = hsegs[bi * (size_t)ncols + (size_t)c];
                sg.out_row0 = rows;
                segs[(size_t)c].push_back(sg);
                if (sch.kind[c] == 3) {
                    const int64_t sum = hbytes[bi * (size_t)ncols + (size_t)c], at = sg.values_off + 4 * sg.n;
                    for (int64_t o = 0; o < sum; o += 1 << 20) copies[(size_t)c].push_back(DeserCopy{at + o, col_bytes[(size_t)c] + o, std::min<int64_t>(1 << 20, sum - o)});
                    col_bytes[(size_t)c] += sum;
                }
            }
            rows += hrows[bi];
            AURON_CHECK(rows < (int64_t)INT32_MAX, "shuffle read: chunk too large");
        }
        metrics.add("size", payload_bytes);
        auto out = std::make_shared<Batch>();
        out->num_rows = rows;
        for (int c = 0; c < ncols; c++)
            out->cols.push_back(deserialize_column(ctx, out_schema.fields[(size_t)c].type, P<uint8_t>(dpayload), segs[(size_t)c], rows, copies[(size_t)c], col_bytes[(size_t)c]));
        ctx.sync();   // dcomp / dpayload are released when this function returns
        metrics.add("output_rows", rows);
        return out;
    }

    // ipc_reader_exec.rs:211-219: a block that cannot be decoded is reported to the host as a fetch failure before the task fails
    BatchPtr next(Task& t) override {
        try {
            return next_chunk(t);
        } catch (const Error& e) {
            const std::string msg = e.what();
            if (msg.rfind("shuffle read:", 0) == 0 && t.cb && t.cb->fetch_failed) t.cb->fetch_failed(t.cb->user, resource_id.c_str(), msg.c_str());
            throw;
        }
    }
    BatchPtr next_chunk(Task& t) {
        if (done) return nullptr;
        OpTimer timer(metrics, "elapsed_ns");
        const int ncols = (int)out_schema.fields.size();
        // ---- 1. gather blocks for this chunk
        std::vector<Block> blocks;
        int64_t fetched = 0;
        const int64_t kChunkBytes = 256ll << 20;   // compressed bytes per chunk
        while (fetched < kChunkBytes) {
            AURON_CHECK(t.is_running(), "task killed");
            Block b;
            if (!next_block(t, &b)) {
                done = true;
                break;
            }
            if (b.length == 0) continue;
            fetched += b.length;
            blocks.push_back(std::move(b));
        }
        if (blocks.empty()) return nullptr;
        // ---- 2. bytes of every block -> one pinned buffer (recycled: no page faults, no frees), split into codec streams
        using Stream = StreamRef;
        std::vector<Stream> streams;
        std::vector<int64_t> boff(blocks.size() + 1, 0);
        for (size_t i = 0; i < blocks.size(); i++) boff[i + 1] = boff[i] + ((blocks[i].length + 63) & ~(int64_t)63);
        size_t ccap = 0;
        uint8_t* cbuf = (uint8_t*)pinned_pool().get((size_t)boff.back() + 64, &ccap);
        struct CGuard {
            uint8_t* p;
            size_t cap;
            ~CGuard() { pinned_pool().put(p, cap); }
        } cguard{cbuf, ccap};
        {
            OpTimer tf(metrics, "fetch_ns");
            parallel_for(blocks.size(), 16, [&](size_t i) {
                Block& b = blocks[i];
                uint8_t* dst = cbuf + boff[i];
                if (b.data) {
                    memcpy(dst, b.data, (size_t)b.length);
                } else {
                    int fd = open(b.path.c_str(), O_RDONLY);
                    AURON_CHECK(fd >= 0, "cannot open shuffle file " + b.path);
                    int64_t got = 0;
                    while (got < b.length) {
                        ssize_t r = pread(fd, dst + got, (size_t)(b.length - got), b.offset + got);
                        if (r <= 0) {
                            close(fd);
                            fail("short read on shuffle file " + b.path);
                        }
                        got += r;
                    }
                    close(fd);
                }
                b.data = dst;
                std::vector<uint8_t>().swap(b.copy);
            });
            for (auto& b : blocks) {
                int64_t pos = 0;
                while (pos < b.length) {
                    AURON_CHECK(pos + 4 <= b.length, "shuffle read: truncated block header");
                    uint32_t len;
                    memcpy(&len, b.data + pos, 4);
                    pos += 4;
                    AURON_CHECK(pos + (int64_t)len <= b.length, "shuffle read: block overruns its segment");
                    if (len) streams.push_back(Stream{b.data + pos, (size_t)len});
                    pos += len;
                }
            }
        }
        if (BatchPtr dev = try_device_decode(t, streams, cbuf, (size_t)boff.back())) return dev;
        // ---- 3. decompress on the worker pool, straight into ONE pinned payload buffer: every stream gets a slot sized by
        // its decoded-size bound (LZ4 frames: blocks x block size; ZSTD: the frame's content size); streams that do not
        // announce a size are decoded into vectors first.  (Decoding into fresh vectors and concatenating cost 540 of
        // 630 ms for 1.8 GB: page faults of 32 threads on one address space.)
        const size_t ns = streams.size();
        std::vector<int64_t> bound(ns), slot(ns + 1, 0), size(ns, 0);
        std::vector<std::vector<uint8_t>> loose(ns);
        {
            OpTimer td(metrics, "decompress_ns");
            parallel_for(ns, 32, [&](size_t i) {
                bound[i] = stream_bound(streams[i].p, streams[i].n);
                if (bound[i] < 0) {
                    decompress_stream(streams[i].p, streams[i].n, loose[i]);
                    bound[i] = (int64_t)loose[i].size();
                }
            });
        }
        for (size_t i = 0; i < ns; i++) slot[i + 1] = slot[i] + ((bound[i] + 63) & ~(int64_t)63);
        int64_t total = slot[ns];
        size_t cap = 0;
        uint8_t* payload = (uint8_t*)pinned_pool().get((size_t)total + 64, &cap);
        struct PinnedGuard {
            uint8_t* p;
            size_t cap;
            ~PinnedGuard() { pinned_pool().put(p, cap); }
        } guard{payload, cap};
        {
            OpTimer td(metrics, "decompress_ns");
            parallel_for(ns, 32, [&](size_t i) {
                if (!loose[i].empty() || bound[i] == 0) {
                    if (!loose[i].empty()) memcpy(payload + slot[i], loose[i].data(), loose[i].size());
                    size[i] = (int64_t)loose[i].size();
                } else size[i] = decompress_stream_into(streams[i].p, streams[i].n, payload + slot[i], bound[i]);
            });
        }
        loose.clear();
        int64_t payload_bytes = 0;
        for (size_t i = 0; i < ns; i++) payload_bytes += size[i];
        metrics.add("size", payload_bytes);
        // ---- 4. layout walk.  The payload of a segment is ONE byte stream (the reference's reader chains blocks), stored
        // here as one slot per codec stream; a batch section that would straddle two slots makes the walk start over on a
        // compacted copy (never happens with the reference's or this engine's writer: blocks hold whole batches).
        std::vector<std::vector<DeserSeg>> segs((size_t)ncols);
        std::vector<std::vector<DeserCopy>> copies((size_t)ncols);
        std::vector<int64_t> col_bytes((size_t)ncols, 0);
        int64_t rows = 0;
        struct Straddle {};
        auto walk = [&]() {
            for (auto& v : segs) v.clear();
            for (auto& v : copies) v.clear();
            std::fill(col_bytes.begin(), col_bytes.end(), 0);
            rows = 0;
            size_t si = 0;
            int64_t pos = slot[0], end = slot[0] + (ns ? size[0] : 0);
            auto hop = [&]() {   // at the end of a slot: continue in the next non-empty one; false at the end of the payload
                while (pos == end) {
                    if (++si >= ns) return false;
                    pos = slot[si];
                    end = slot[si] + size[si];
                }
                return true;
            };
            auto section = [&](int64_t len) {   // [pos, pos+len) must be contiguous
                if (len == 0) return pos;
                if (!hop()) fail("shuffle read: batch overruns the payload");
                if (pos + len > end) throw Straddle();
                int64_t at = pos;
                pos += len;
                return at;
            };
            auto varint = [&]() {
                uint64_t v = 0;
                for (int shift = 0;; shift += 7) {
                    AURON_CHECK(shift < 64 && hop(), "shuffle read: truncated varint");
                    uint8_t b = payload[pos++];
                    v |= (uint64_t)(b & 0x7f) << shift;
                    if (!(b & 0x80)) return v;
                }
            };
            if (ns == 0) return;
            while (hop()) {
                const int64_t n = (int64_t)varint();
                for (int c = 0; c < ncols; c++) {
                    const DType& ty = out_schema.fields[(size_t)c].type;
                    if (ty.id == T_NULL) continue;
                    DeserSeg sg{-1, 0, rows, n};
                    if (varint()) sg.validity_off = section((n + 7) / 8);
                    if (ty.id == T_BOOL) sg.values_off = section((n + 7) / 8);
                    else if (ty.is_varlen()) {
                        sg.values_off = section(4 * n);
                        const uint8_t *p0 = payload + sg.values_off, *p1 = p0 + n, *p2 = p1 + n, *p3 = p2 + n;
                        int64_t sum = 0;
                        for (int64_t i = 0; i < n; i++) sum += (int64_t)((uint32_t)p0[i] | ((uint32_t)p1[i] << 8) | ((uint32_t)p2[i] << 16) | ((uint32_t)p3[i] << 24));
                        const int64_t at = section(sum);
                        for (int64_t o = 0; o < sum; o += 1 << 20)
                            copies[(size_t)c].push_back(DeserCopy{at + o, col_bytes[(size_t)c] + o, std::min<int64_t>(1 << 20, sum - o)});
                        col_bytes[(size_t)c] += sum;
                    } else sg.values_off = section((int64_t)ty.width() * n);
                    segs[(size_t)c].push_back(sg);
                }
                rows += n;
                AURON_CHECK(rows < (int64_t)INT32_MAX, "shuffle read: chunk too large");
            }
        };
        {
            OpTimer tw(metrics, "layout_ns");
            try {
                walk();
            } catch (const Straddle&) {
                int64_t w = 0;
                for (size_t i = 0; i < ns; i++) {   // compact: slots are in payload order, so moving left is safe
                    memmove(payload + w, payload + slot[i], (size_t)size[i]);
                    w += size[i];
                }
                for (size_t i = 0; i < ns; i++) {
                    slot[i] = i == 0 ? 0 : 0;
                    size[i] = i == 0 ? w : 0;
                }
                total = w;
                walk();
            }
        }
        // ---- 5. device
        OpTimer tdev(metrics, "device_ns");
        Buf dpayload = dalloc(t.ctx, (size_t)total + 64);
        CUDA_OK(cudaMemcpyAsync(dpayload->ptr, payload, (size_t)total, cudaMemcpyHostToDevice, t.ctx.stream));
        auto out = std::make_shared<Batch>();
        out->num_rows = rows;
        for (int c = 0; c < ncols; c++)
            out->cols.push_back(deserialize_column(t.ctx, out_schema.fields[(size_t)c].type, P<uint8_t>(dpayload), segs[(size_t)c], rows, copies[(size_t)c], col_bytes[(size_t)c]));
        t.ctx.sync();   // the pinned payload goes back to the pool
        metrics.add("output_rows", rows);
        return out;
    }
};

OperatorPtr make_ipc_reader(Task&, const Schema& schema, const std::string& resource_id) {
    auto op = std::make_unique<IpcReaderExec>();
    op->name = "IpcReaderExec";
    op->out_schema = schema;
    op->resource_id = resource_id;
    return op;
}

}  // namespace auron
