// exchange.cu -- the one data-path collective of the hot path: the hash repartition between
// stages.  In the reference this is Spark's sort-shuffle (files + block manager; datafusion-ext-plans/src/shuffle/*,
// ipc_reader_exec.rs).  Inside one 8-GPU box the same repartition is: murmur3 partition ids -> counting sort ->
// partition-contiguous gather (all on device, k_hash.cu / k_sort.cu / k_basic.cu) followed by a variable-size
// all-to-all over NVLink 5 / NVSwitch: counts are all-gathered first, then every column buffer moves with grouped
// ncclSend / ncclRecv (one group per buffer kind), validity travels as one byte per row and is re-packed on arrival,
// utf8 travels as lengths + bytes and offsets are rebuilt by a prefix scan.  Partition p is owned by rank
// p * world / num_parts (contiguous blocks, so rows sorted by partition are already sorted by destination rank).
//
// NCCL is loaded with dlopen (the torch-bundled libnccl.so.2 / system libnccl): no link-time dependency, and the
// product never falls back to a host path -- without NCCL the exchange fails loudly.
#include <dlfcn.h>

#include <mutex>

#include "device_utils.cuh"
#include "exchange.h"
#include "kernels.h"

namespace auron {

#define LAUNCH_CHECK(ctx)            \
    do {                             \
        CUDA_OK(cudaGetLastError()); \
        launch_count(ctx);           \
    } while (0)

// ---------------------------------------------------------------------------------------------- NCCL via dlopen
typedef void* ncclComm_t;
struct NcclUniqueId {
    char internal[128];
};
enum { NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_INT64 = 4 };
struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        if (const char* p = getenv("AURON_NCCL_LIB")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
#define SYM(field, name) api.field = (decltype(api.field))dlsym(h, name)
        SYM(GetUniqueId, "ncclGetUniqueId");
        SYM(CommInitRank, "ncclCommInitRank");
        SYM(CommDestroy, "ncclCommDestroy");
        SYM(AllGather, "ncclAllGather");
        SYM(Send, "ncclSend");
        SYM(Recv, "ncclRecv");
        SYM(GroupStart, "ncclGroupStart");
        SYM(GroupEnd, "ncclGroupEnd");
        SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    });
    return api;
}
#define NCCL_OK(expr)                                                                                                   \
    do {                                                                                                                \
        int _r = (expr);                                                                                                \
        if (_r != 0) fail(std::string("NCCL error: ") + (nccl().GetErrorString ? nccl().GetErrorString(_r) : "?") + " in " #expr); \
    } while (0)

static struct {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
} g_comm;

void nccl_get_unique_id(uint8_t out[128]) {
    AURON_CHECK(nccl().ok, "NCCL library not found (set AURON_NCCL_LIB to libnccl.so.2)");
    NcclUniqueId id;
    NCCL_OK(nccl().GetUniqueId(&id));
    memcpy(out, id.internal, 128);
}
void nccl_init(const uint8_t id_bytes[128], int rank, int world, int device) {
    AURON_CHECK(nccl().ok, "NCCL library not found (set AURON_NCCL_LIB to libnccl.so.2)");
    AURON_CHECK(g_comm.comm == nullptr, "NCCL communicator already initialised");
    CUDA_OK(cudaSetDevice(device));
    NcclUniqueId id;
    memcpy(id.internal, id_bytes, 128);
    NCCL_OK(nccl().CommInitRank(&g_comm.comm, world, id, rank));
    g_comm.rank = rank;
    g_comm.world = world;
    g_comm.device = device;
}
void nccl_finalize() {
    if (g_comm.comm && nccl().CommDestroy) nccl().CommDestroy(g_comm.comm);
    g_comm.comm = nullptr;
    g_comm.world = 1;
    g_comm.rank = 0;
}
int nccl_world() { return g_comm.comm ? g_comm.world : 1; }
int nccl_rank() { return g_comm.rank; }

// ---------------------------------------------------------------------------------------------- small kernels
__global__ void bits_to_bytes_kernel(const uint8_t* __restrict__ bits, int64_t n, uint8_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = bits ? (uint8_t)bit_get(bits, i) : (uint8_t)1;
}
__global__ void __launch_bounds__(256) bytes_to_bits_kernel(const uint8_t* __restrict__ bytes, int64_t n, uint32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    bool b = i < n && bytes[i] != 0;
    uint32_t w = __ballot_sync(FULL_MASK, b);
    if (lane_id() == 0 && i < n) out[i >> 5] = w;
}
__global__ void lens_from_offsets_kernel(const int32_t* __restrict__ off, int64_t n, int32_t* __restrict__ lens) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lens[i] = off[i + 1] - off[i];
}

// grouped variable-size exchange of one buffer: element size `esz`, send_counts / recv_counts in elements
static Buf exchange_buffer(Ctx& ctx, const uint8_t* send, int esz, const std::vector<int64_t>& send_off, const std::vector<int64_t>& send_cnt,
                           const std::vector<int64_t>& recv_off, const std::vector<int64_t>& recv_cnt) {
    int world = g_comm.world;
    int64_t total = recv_off[world - 1] + recv_cnt[world - 1];
    Buf out = dalloc(ctx, (size_t)std::max<int64_t>(total, 1) * esz);
    NCCL_OK(nccl().GroupStart());
    for (int r = 0; r < world; r++) {
        if (send_cnt[r]) NCCL_OK(nccl().Send(send + send_off[r] * esz, (size_t)(send_cnt[r] * esz), NCCL_UINT8, r, g_comm.comm, ctx.stream));
        if (recv_cnt[r]) NCCL_OK(nccl().Recv(P<uint8_t>(out) + recv_off[r] * esz, (size_t)(recv_cnt[r] * esz), NCCL_UINT8, r, g_comm.comm, ctx.stream));
    }
    NCCL_OK(nccl().GroupEnd());
    return out;
}

// num_parts > 0: rows [part_row_off[p], part_row_off[p + 1]) go to the rank that owns partition p (all-to-all-v).
// num_parts == 0: every rank receives every rank's rows, in rank order (all-gather-v) -- the build side of a broadcast join, collected
// from the partitions the ranks hold and replicated over NVLink instead of once per GPU over PCIe.
BatchPtr nccl_exchange(Ctx& ctx, const Batch& sorted, const std::vector<int64_t>& part_row_off, int64_t num_parts, int64_t* bytes_sent) {
    AURON_CHECK(g_comm.comm != nullptr, "NCCL exchange requested but auron_b200_nccl_init was not called");
    AURON_CHECK(ctx.device == g_comm.device, "exchange on a different device than the communicator");
    const int world = g_comm.world;
    // rows per destination rank: partition p belongs to rank p * world / num_parts
    std::vector<int64_t> send_off(world + 1, 0);   // destination r gets rows [send_off[r], send_off[r] + send_cnt[r])
    std::vector<int64_t> send_cnt(world);
    if (num_parts > 0) {
        int r = 0;
        for (int64_t p = 0; p <= num_parts; p++) {
            int owner = p == num_parts ? world : (int)(p * world / num_parts);
            while (r < owner) send_off[++r] = part_row_off[p];
        }
        for (int q = 0; q < world; q++) send_cnt[q] = send_off[q + 1] - send_off[q];
    } else {
        for (int q = 0; q < world; q++) send_cnt[q] = sorted.num_rows;
    }
    const size_t ncols = sorted.cols.size();
    // counts matrix: [rows, bytes of each varlen column] per destination
    std::vector<int> varlen_cols;
    for (size_t c = 0; c < ncols; c++)
        if (sorted.cols[c]->type.is_varlen()) varlen_cols.push_back((int)c);
    const int kstride = 1 + (int)varlen_cols.size();
    std::vector<int64_t> my_counts((size_t)world * kstride, 0);
    std::vector<std::vector<int64_t>> byte_off(varlen_cols.size());
    for (size_t v = 0; v < varlen_cols.size(); v++) {
        const Column& col = *sorted.cols[varlen_cols[v]];
        std::vector<int32_t> tmp(2 * (size_t)world);   // byte offsets at the first row and behind the last row of every destination's range
        for (int r = 0; r < world; r++) {
            to_host(ctx, &tmp[2 * (size_t)r], P<int32_t>(col.offsets) + send_off[r], 4);
            to_host(ctx, &tmp[2 * (size_t)r + 1], P<int32_t>(col.offsets) + send_off[r] + send_cnt[r], 4);
        }
        byte_off[v].assign(tmp.begin(), tmp.end());
    }
    for (int r = 0; r < world; r++) {
        my_counts[(size_t)r * kstride] = send_cnt[r];
        for (size_t v = 0; v < varlen_cols.size(); v++) my_counts[(size_t)r * kstride + 1 + v] = byte_off[v][2 * (size_t)r + 1] - byte_off[v][2 * (size_t)r];
    }
    Buf d_my = to_device(ctx, my_counts.data(), my_counts.size() * 8);
    Buf d_all = dalloc(ctx, (size_t)world * my_counts.size() * 8);
    NCCL_OK(nccl().AllGather(d_my->ptr, d_all->ptr, my_counts.size(), NCCL_INT64, g_comm.comm, ctx.stream));
    std::vector<int64_t> all((size_t)world * my_counts.size());
    to_host(ctx, all.data(), d_all->ptr, all.size() * 8);
    // what I receive from rank s: all[s][me]
    auto recv_counts = [&](int k) {
        std::vector<int64_t> cnt(world), off(world);
        int64_t acc = 0;
        for (int s = 0; s < world; s++) {
            cnt[s] = all[(size_t)s * my_counts.size() + (size_t)g_comm.rank * kstride + k];
            off[s] = acc;
            acc += cnt[s];
        }
        return std::make_pair(off, cnt);
    };
    auto rrows = recv_counts(0);
    const int64_t n_recv = rrows.first[world - 1] + rrows.second[world - 1];
    std::vector<int64_t> send_off_rows(send_off.begin(), send_off.begin() + world);
    auto out = std::make_shared<Batch>();
    out->num_rows = n_recv;
    int64_t sent = 0;
    const int64_t n_rows = sorted.num_rows;
    unsigned rb = (unsigned)((n_rows + 255) / 256), ob = (unsigned)((n_recv + 255) / 256);
    for (size_t c = 0; c < ncols; c++) {
        const Column& col = *sorted.cols[c];
        auto oc = std::make_shared<Column>();
        oc->type = col.type;
        oc->len = n_recv;
        if (col.type.id == T_NULL) {
            oc->null_count = n_recv;
            out->cols.push_back(oc);
            continue;
        }
        // validity: every rank must take the same path, so it is always exchanged (one byte per row)
        {
            Buf vb = dalloc(ctx, (size_t)std::max<int64_t>(n_rows, 1));
            if (n_rows) {
                bits_to_bytes_kernel<<<rb, 256, 0, ctx.stream>>>(col.vbits(), n_rows, P<uint8_t>(vb));
                LAUNCH_CHECK(ctx);
            }
            Buf rv = exchange_buffer(ctx, P<uint8_t>(vb), 1, send_off_rows, send_cnt, rrows.first, rrows.second);
            oc->validity = dalloc(ctx, bitmap_alloc_bytes(n_recv));
            oc->null_count = -1;
            if (n_recv) {
                bytes_to_bits_kernel<<<ob, 256, 0, ctx.stream>>>(P<uint8_t>(rv), n_recv, P<uint32_t>(oc->validity));
                LAUNCH_CHECK(ctx);
            }
            sent += n_rows;
        }
        if (col.type.id == T_BOOL) {
            Buf vb = dalloc(ctx, (size_t)std::max<int64_t>(n_rows, 1));
            if (n_rows) {
                bits_to_bytes_kernel<<<rb, 256, 0, ctx.stream>>>(P<uint8_t>(col.data), n_rows, P<uint8_t>(vb));
                LAUNCH_CHECK(ctx);
            }
            Buf rv = exchange_buffer(ctx, P<uint8_t>(vb), 1, send_off_rows, send_cnt, rrows.first, rrows.second);
            oc->data = dalloc(ctx, bitmap_alloc_bytes(n_recv));
            if (n_recv) {
                bytes_to_bits_kernel<<<ob, 256, 0, ctx.stream>>>(P<uint8_t>(rv), n_recv, P<uint32_t>(oc->data));
                LAUNCH_CHECK(ctx);
            }
            sent += n_rows;
        } else if (col.type.width() > 0) {
            oc->data = exchange_buffer(ctx, P<uint8_t>(col.data), col.type.width(), send_off_rows, send_cnt, rrows.first, rrows.second);
            sent += n_rows * col.type.width();
        } else {   // utf8 / binary: lengths, then bytes
            size_t v = 0;
            while (varlen_cols[v] != (int)c) v++;
            Buf lens = dalloc(ctx, (size_t)std::max<int64_t>(n_rows, 1) * 4);
            if (n_rows) {
                lens_from_offsets_kernel<<<rb, 256, 0, ctx.stream>>>(P<int32_t>(col.offsets), n_rows, P<int32_t>(lens));
                LAUNCH_CHECK(ctx);
            }
            Buf rl = exchange_buffer(ctx, P<uint8_t>(lens), 4, send_off_rows, send_cnt, rrows.first, rrows.second);
            oc->offsets = dalloc(ctx, (size_t)(n_recv + 1) * 4);
            exclusive_scan_i32(ctx, P<int32_t>(rl), P<int32_t>(oc->offsets), n_recv, P<int32_t>(oc->offsets) + n_recv);
            if (n_recv == 0) CUDA_OK(cudaMemsetAsync(oc->offsets->ptr, 0, 4, ctx.stream));
            auto rbytes = recv_counts(1 + (int)v);
            std::vector<int64_t> sboff(world), sbcnt(world);
            for (int r = 0; r < world; r++) {
                sboff[r] = byte_off[v][2 * (size_t)r];
                sbcnt[r] = byte_off[v][2 * (size_t)r + 1] - byte_off[v][2 * (size_t)r];
            }
            oc->data = exchange_buffer(ctx, P<uint8_t>(col.data), 1, sboff, sbcnt, rbytes.first, rbytes.second);
            oc->data_bytes = rbytes.first[world - 1] + rbytes.second[world - 1];
            AURON_CHECK(oc->data_bytes <= (int64_t)INT32_MAX, "utf8 column exceeds 2 GiB after the exchange");
            sent += n_rows * 4 + col.data_bytes;
        }
        out->cols.push_back(oc);
    }
    ctx.sync();
    if (bytes_sent) *bytes_sent = sent * (num_parts > 0 ? 1 : world);   // (the all-gather sends every row to every rank)
    return out;
}

}  // namespace auron
