// C ABI of libspumoni_b200.so (declared in include/spumoni_b200.h).
//
// Host plumbing only: index lifetime, device residency (one or several GPUs), batch pipelining (H2D / kernels / D2H on
// rotating streams, one host thread per GPU) and error translation.  All per-base work happens in kernels.cu.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <deque>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/spumoni_b200.h"
#include "flat_index.h"
#include "build_kernels.h"
#include "host_util.h"
#include "index_host.h"
#include "kernels.h"

namespace {

thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
#define SPB_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(SPB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
    } while (0)

// Host entry points: the batch is cut into sub-batches that rotate over `n_slots` streams.  One read is one thread, so
// what keeps the ~170 000 resident lanes of a B200 busy is the number of READS in flight: a sub-batch is sized by reads
// (kWantReads, so that the slots together hold about two waves), bounded below and above in bases (copy efficiency, device
// buffers), and several sub-batches are in flight (copies of some overlapping the kernels of others).
constexpr int kMaxSlots = 16;
constexpr int kDefaultSlots = 8;                       // SPUMONI_B200_SLOTS
constexpr uint64_t kWantReads = 43000;                 // SPUMONI_B200_CHUNK_READS
constexpr uint64_t kMinChunkBases = 8ull << 20;        // SPUMONI_B200_CHUNK_MB (in Mi bases): lower bound of a sub-batch that has enough reads
constexpr uint64_t kMaxChunkBases = 256ull << 20;      // upper bound (device buffers: 8 slots x (1 + 4 [+ 8 + 4]) bytes per base)
constexpr int kCounters = 2048;   // work-counter pairs are rotated per launch: at most kCounters / 2 launches of one handle in flight
constexpr int kDefaultLanes = 4;  // cooperative kernel (indexes without fast symbols)

// a device buffer that only ever grows; the owner's stream must be idle when it does
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    int reserve(size_t n, size_t slack = 0) {
        if (n <= cap) return SPB_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        SPB_CUDA(cudaMalloc(&p, (n + slack) * sizeof(T)));
        cap = n;
        return SPB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return SPB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        SPB_CUDA(cudaMallocHost(&p, n * sizeof(T)));
        cap = n;
        return SPB_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Slot {
    cudaStream_t stream = nullptr;
    DevBuf<uint8_t> d_reads, d_out;      // d_out: lengths at the requested width, or the bit map
    DevBuf<uint64_t> d_offsets, d_pointers_raw;  // pointers: bytes / 8
    DevBuf<uint32_t> d_docs, d_order, d_packed;
    DevBuf<spb_bin_stats> d_stats;
    DevBuf<uint64_t> d_exc_pos;
    DevBuf<uint8_t> d_exc_val;
    PinBuf<uint64_t> h_offsets;
    PinBuf<uint32_t> h_order;
    unsigned long long* d_sat = nullptr;  // clamped-value counter of the launch in flight
    unsigned long long* h_sat = nullptr;  // pinned copy
    bool sat_pending = false;
    // on-device text formatting (spb_pml_query_host_text)
    DevBuf<uint8_t> d_ids;
    DevBuf<uint64_t> d_id_offsets;
    PinBuf<uint64_t> h_id_offsets;
    struct TextStream {  // one formatted output: lengths, MS pointers, document numbers
        DevBuf<uint64_t> d_starts;
        DevBuf<char> d_text;
    } ts[3];
    uint64_t* h_total = nullptr;  // 3 x uint64, pinned
    void release() {
        d_reads.release(); d_out.release(); d_offsets.release(); d_pointers_raw.release(); d_docs.release(); d_order.release();
        d_packed.release(); d_stats.release(); d_exc_pos.release(); d_exc_val.release(); h_offsets.release(); h_order.release();
        d_ids.release(); d_id_offsets.release(); h_id_offsets.release();
        for (auto& t : ts) { t.d_starts.release(); t.d_text.release(); }
        if (d_sat) cudaFree(d_sat);
        if (h_sat) cudaFreeHost(h_sat);
        if (h_total) cudaFreeHost(h_total);
        d_sat = nullptr; h_sat = nullptr; h_total = nullptr; sat_pending = false;
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};

// The persisting-L2 carve-out is a device-wide limit of the process and it costs every kernel that does not use it (MS on
// config 2: 28.6 instead of 31.9 Gbases/s with the carve-out set): it is held only while at least one index that profits from
// it (PML, row table up to twice the carve-out) is resident on the device, and given back with the last one.
std::mutex g_persist_mutex;
int g_persist_users[64] = {0};

// everything one GPU needs to serve queries on a resident copy of the flat image
struct DeviceCtx {
    int device = -1;
    const uint8_t* dev_blob = nullptr;
    bool own_dev = false;
    int num_sms = 0;
    unsigned long long* d_counters = nullptr;
    std::atomic<uint32_t> counter_next{0};
    Slot slots[kMaxSlots];
    PinBuf<uint8_t> h_edges;  // bit-map transport: first / last 32-byte block of every sub-batch (merged by the host)
    uint64_t persist_size = 0, persist_window = 0;  // SPUMONI_B200_L2_PERSIST: persisting L2 carve-out / largest access-policy window
    ~DeviceCtx() {
        if (device < 0) return;
        cudaSetDevice(device);
        for (auto& s : slots) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            s.release();
        }
        h_edges.release();
        if (persist_size && device < 64) {  // give the L2 back with the last user
            std::lock_guard<std::mutex> lk(g_persist_mutex);
            if (--g_persist_users[device] == 0) {
                cudaCtxResetPersistingL2Cache();
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
                cudaGetLastError();
            }
        }
        if (d_counters) cudaFree(d_counters);
        if (own_dev && dev_blob) cudaFree((void*)dev_blob);
    }
};

}  // namespace

struct spb_index {
    int mode = 0;
    SpbFlatHeader hdr;
    spb::FlatBlob host_blob;
    std::vector<std::unique_ptr<DeviceCtx>> devs;  // devs[0] serves the *_device entry points
    uint32_t row_ld = 1;  // full (32-byte) rows allocate in L1 like the others; SPUMONI_B200_ROW_LD=0 makes them bypass it
    int n_slots = kDefaultSlots;
    uint64_t want_reads = kWantReads, min_chunk_bases = kMinChunkBases, max_chunk_bases = kMaxChunkBases;
    bool configured = false;
};

namespace {

void configure(spb_index* idx) {
    if (idx->configured) return;
    idx->configured = true;
    if (const char* e = std::getenv("SPUMONI_B200_ROW_LD")) idx->row_ld = (uint32_t)std::atoi(e);  // development knob
    if (const char* e = std::getenv("SPUMONI_B200_SLOTS")) { int v = std::atoi(e); if (v >= 1 && v <= kMaxSlots) idx->n_slots = v; }
    if (const char* e = std::getenv("SPUMONI_B200_CHUNK_MB")) { long v = std::atol(e); if (v >= 1 && v <= 4096) idx->min_chunk_bases = (uint64_t)v << 20; }
    if (const char* e = std::getenv("SPUMONI_B200_CHUNK_READS")) { long v = std::atol(e); if (v >= 1) idx->want_reads = (uint64_t)v; }
    if (const char* e = std::getenv("SPUMONI_B200_TEST_CHUNK_BASES")) { long v = std::atol(e); if (v >= 1) idx->min_chunk_bases = idx->max_chunk_bases = (uint64_t)v; }  // tests: tiny sub-batches
    if (idx->max_chunk_bases < idx->min_chunk_bases) idx->max_chunk_bases = idx->min_chunk_bases;
}

int ensure_ctx(spb_index* idx, DeviceCtx& c) {
    configure(idx);
    SPB_CUDA(cudaSetDevice(c.device));
    if (!c.d_counters) {
        SPB_CUDA(cudaMalloc(&c.d_counters, kCounters * sizeof(unsigned long long)));
        SPB_CUDA(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, c.device));
        // Opt-in only (it is a process-wide device setting): SPUMONI_B200_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity.
        // Measured without effect on the gathers of this path (profiles/r1_gather_bench_exp1.txt), so the default leaves it alone.
        // PML over a row table that (nearly) fits the L2: the rows get the persisting part of the L2 through an access-policy
        // window on the walk launches (the read / output streams are evict-first already): +4-6 % on config 2
        // (profiles/r2i_experiments.json: 84.9 against 81.3 Gbases/s; a 60 % carve-out and MS, whose 16-byte slots are twice the
        // L2's effective size, lose).  cudaLimitPersistingL2CacheSize is a device-wide setting of this process: it is set to
        // the maximum while an index that uses it is resident and back to zero when the last one closes;
        // SPUMONI_B200_L2_PERSIST=0 leaves it alone (=<percent of the maximum> for experiments).
        {
            int pct = 100;
            if (const char* e = std::getenv("SPUMONI_B200_L2_PERSIST")) pct = std::atoi(e);
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c.device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c.device);
            const SpbFlatHeader& h = idx->hdr;
            const uint64_t row_bytes = h.off_mrows8 ? h.mrows8_stride : (h.off_mrows32 ? 32u : 16u);
            const uint64_t tbytes = (h.R + SPB_MV_ROW_PAD) * row_bytes;
            const size_t want = (size_t)((double)max_persist * std::min(std::max(pct, 0), 100) / 100.0);
            const bool eligible = idx->mode == SPB_MODE_PML && h.num_fast && pct > 0 && max_persist > 0 && tbytes <= 2 * (uint64_t)want && tbytes <= (uint64_t)max_window;
            if (eligible && c.device < 64) {
                std::lock_guard<std::mutex> lk(g_persist_mutex);
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                    g_persist_users[c.device]++;
                    c.persist_size = want;
                    c.persist_window = (uint64_t)max_window;
                }
                cudaGetLastError();
                if (std::getenv("SPUMONI_B200_TIMING")) std::fprintf(stderr, "[timing] L2 persisting carve-out %.1f MB (max %.1f) for a %.1f MB row table\n", want / 1e6, max_persist / 1e6, tbytes / 1e6);
            }
        }
        if (const char* e = std::getenv("SPUMONI_B200_L2_FETCH")) {
            const size_t gran = (size_t)std::atoi(e);
            if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);  // a hint: failure is not an error
            cudaGetLastError();
        }
    }
    return SPB_OK;
}

DeviceCtx* primary(spb_index* idx) { return idx->devs.empty() ? nullptr : idx->devs[0].get(); }

int ensure_device_state(spb_index* idx) {
    DeviceCtx* c = primary(idx);
    if (!c || !c->dev_blob) return fail(SPB_ERR_NO_DEVICE, "index is not resident on a GPU (call spb_index_to_device)");
    return ensure_ctx(idx, *c);
}

int ensure_slot_base(Slot& s) {
    if (!s.stream) SPB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (!s.d_sat) SPB_CUDA(cudaMalloc(&s.d_sat, sizeof(unsigned long long)));
    if (!s.h_sat) SPB_CUDA(cudaMallocHost(&s.h_sat, sizeof(unsigned long long)));
    return SPB_OK;
}

// The host entry points hand caller memory to asynchronous copies: whatever way they return (errors included), no
// copy may still be in flight afterwards, and no clamped-value count of this call may leak into the next one.
struct DrainOnExit {
    spb_index* idx;
    ~DrainOnExit() {
        for (auto& c : idx->devs) {
            cudaSetDevice(c->device);
            for (auto& s : c->slots) {
                if (s.stream) cudaStreamSynchronize(s.stream);
                s.sat_pending = false;
            }
        }
    }
};

// SPB_LENGTH_BITS -> 0 (kernels' out_width); 0 -> 4
int out_width_of(const spb_query_opts* opts) {
    const uint32_t w = opts ? opts->length_bytes : 0;
    if (w == SPB_LENGTH_BITS) return 0;
    return w == 0 ? 4 : (int)w;
}
bool width_ok(const spb_query_opts* opts) {
    const uint32_t w = opts ? opts->length_bytes : 0;
    return w == 0 || w == 1 || w == 2 || w == 4 || w == SPB_LENGTH_BITS;
}

struct LaunchExtras {
    const uint32_t* d_order = nullptr;
    uint64_t bit_shift = 0;
};

int launch_query(spb_index* idx, DeviceCtx& c, bool ms, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                 void* d_lengths, uint64_t* d_pointers, uint32_t* d_docs, spb_bin_stats* d_stats, const spb_query_opts* opts,
                 unsigned long long* d_saturated, cudaStream_t stream, const LaunchExtras& ex) {
    if (n_reads == 0) return SPB_OK;
    if ((ms ? 1 : 0) != idx->mode) return fail(SPB_ERR_ARG, ms ? "MS query on a PML index" : "PML query on an MS index");
    if (!d_reads || !d_offsets || !d_lengths || (ms && !d_pointers)) return fail(SPB_ERR_ARG, "null device buffer");
    if (((uintptr_t)d_reads & 255u) != 0) return fail(SPB_ERR_ARG, "d_reads must be 256-byte aligned");
    if (((uintptr_t)d_lengths & 15u) != 0 || ((uintptr_t)d_pointers & 15u) != 0) return fail(SPB_ERR_ARG, "d_lengths / d_pointers must be 16-byte aligned");
    const SpbFlatHeader& h = idx->hdr;
    const bool has_fast = h.num_fast > 0 && (h.off_mrows != 0 || h.off_mrows8 != 0 || h.off_mrows32 != 0) && (!ms || h.off_msamp != 0 || h.mrows8_stride == 16);
    int lanes = opts && opts->lanes_per_read ? (int)opts->lanes_per_read : (has_fast ? 1 : kDefaultLanes);
    if (lanes != 1 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) return fail(SPB_ERR_ARG, "lanes_per_read must be 1, 4, 8, 16 or 32");
    if (lanes == 1 && !has_fast) return fail(SPB_ERR_ARG, "lanes_per_read = 1 needs an index with fast symbols (frequent run heads < 0x80)");
    if (!width_ok(opts)) return fail(SPB_ERR_ARG, "length_bytes must be 0, 1, 2, 4 or SPB_LENGTH_BITS");
    const int width = out_width_of(opts);
    if (width != 4 && lanes != 1) return fail(SPB_ERR_ARG, "narrow length output needs lanes_per_read = 1");
    if (width == 0 && (ms || d_docs)) return fail(SPB_ERR_ARG, "SPB_LENGTH_BITS describes PML lengths only (no MS, no document numbers)");
    if (width == 0 && ((uintptr_t)d_lengths & 31u) != 0) return fail(SPB_ERR_ARG, "the bit map must be 32-byte aligned");
    if (d_docs) {
        if (!h.num_docs || !h.off_cdoc) return fail(SPB_ERR_ARG, "document numbers requested but the index was opened without SPB_OPEN_DOC");
        if (lanes != 1 || !h.off_mdoc) return fail(SPB_ERR_ARG, "document numbers need the thread-per-read kernels (lanes_per_read = 1, an index with fast symbols)");
        if (width != 4) return fail(SPB_ERR_ARG, "document numbers need 4-byte lengths");
        if (((uintptr_t)d_docs & 15u) != 0) return fail(SPB_ERR_ARG, "d_doc_nums must be 16-byte aligned");
    }
    spb::WalkParams p;
    std::memset(&p, 0, sizeof(p));
    p.runs = reinterpret_cast<const uint4*>(c.dev_blob + h.off_runs);
    p.blocks = c.dev_blob + h.off_blocks;
    p.crec = reinterpret_cast<const uint4*>(c.dev_blob + h.off_crec);
    p.csamp = ms ? reinterpret_cast<const uint4*>(c.dev_blob + h.off_csamp) : nullptr;
    p.text = ms ? c.dev_blob + h.off_text : nullptr;
    p.chartab = reinterpret_cast<const SpbCharTab*>(c.dev_blob + offsetof(SpbFlatHeader, chartab));
    p.n = h.n; p.R = h.R; p.n_text = h.n_text; p.last_run_sample = h.last_run_sample;
    p.blk_stride = h.blk_stride;
    p.flags = (opts && (opts->flags & SPB_QUERY_TOUPPER)) ? SPB_FLAG_TOUPPER : 0;
    p.ptr32 = (ms && opts && (opts->flags & SPB_QUERY_POINTERS32)) ? 1u : 0u;
    if (p.ptr32 && (lanes != 1 || d_docs)) return fail(SPB_ERR_ARG, "SPB_QUERY_POINTERS32 needs lanes_per_read = 1 and no document numbers");
    p.reads = d_reads; p.offsets = d_offsets; p.n_reads = n_reads;
    p.lengths = d_lengths; p.pointers = d_pointers;
    p.bin_width = (opts && d_stats) ? opts->bin_width : 0;
    p.stats = p.bin_width ? d_stats : nullptr;
    p.max_value_thr = opts ? opts->max_value_thr : 0;
    p.rows = (has_fast && h.off_mrows) ? reinterpret_cast<const uint4*>(c.dev_blob + h.off_mrows) : nullptr;
    p.rows8 = (has_fast && h.off_mrows8) ? reinterpret_cast<const uint64_t*>(c.dev_blob + h.off_mrows8) : nullptr;
    p.rows32 = (has_fast && h.off_mrows32) ? c.dev_blob + h.off_mrows32 : nullptr;
    p.msamp = (has_fast && ms && h.off_msamp) ? reinterpret_cast<const uint4*>(c.dev_blob + h.off_msamp) : nullptr;
    p.rows8_stride = h.mrows8_stride;
    p.msamp_stride = h.msamp_stride;
    for (int i = 0; i < 4; i++) { p.mv_end_k[i] = h.mv_end_k[i]; p.mv_end_off[i] = h.mv_end_off[i]; p.mv_end_sl[i] = h.mv_end_sl[i]; }
    p.out_width = (uint32_t)width;
    p.bit_shift = ex.bit_shift;
    p.order = lanes == 1 ? ex.d_order : nullptr;
    p.row_ld = idx->row_ld;
    p.saturated = d_saturated;
    p.docs = d_docs;
    if (d_docs) {
        p.cdoc = reinterpret_cast<const uint2*>(c.dev_blob + h.off_cdoc);
        p.mdoc = reinterpret_cast<const uint2*>(c.dev_blob + h.off_mdoc);
        p.doc_init = h.doc_init; p.doc_absent_ms = h.doc_absent_ms;
        for (int i = 0; i < 4; i++) p.mv_end_doc[i] = h.mv_end_doc[i];
    }
    if (c.persist_size && lanes == 1 && !ms) {  // (eligibility -- PML, table up to twice the carve-out -- was decided in ensure_ctx)
        const void* tb = p.rows8 ? (const void*)p.rows8 : (p.rows32 ? p.rows32 : (const void*)p.rows);
        const uint64_t row_bytes = p.rows8 ? h.mrows8_stride : (p.rows32 ? 32u : 16u);
        const uint64_t tbytes = (h.R + SPB_MV_ROW_PAD) * row_bytes;
        p.persist_base = tb;
        p.persist_bytes = tbytes;
        p.persist_ratio = (float)std::min(1.0, (double)c.persist_size / (double)tbytes);
    }
    p.work_counter = c.d_counters + 2 * (c.counter_next.fetch_add(1) % (kCounters / 2));
    SPB_CUDA(cudaMemsetAsync(p.work_counter, 0, 2 * sizeof(unsigned long long), stream));
    int launched = 0;
    SPB_CUDA(spb::launch_walk(p, ms, lanes, c.num_sms, stream, &launched));
    g_launches.fetch_add((uint64_t)launched);
    return SPB_OK;
}

// ---- host query ------------------------------------------------------------------------------------------------------
struct HostJob {
    bool ms = false;
    const uint8_t* reads = nullptr;           // one byte per base, or
    const uint8_t* packed = nullptr;          // 2 bits per base + exceptions
    const uint64_t* exc_pos = nullptr;
    const uint8_t* exc_val = nullptr;
    uint64_t n_exc = 0;
    const uint64_t* offsets = nullptr;
    void* lengths = nullptr;
    void* pointers = nullptr;
    uint32_t* docs = nullptr;
    spb_bin_stats* stats = nullptr;
    const spb_query_opts* opts = nullptr;
    int width = 4;          // 4 / 2 / 1 bytes, 0 = bits
    bool ptr32 = false, want_stats = false, count_sat = false;
};

struct EdgeRec { uint64_t first_blk, n_blk; };  // bit-map blocks of a sub-batch (32 bytes each), global block indices

// All sub-batches chunks[c_begin, c_end) on one GPU.  edges: 64 bytes per chunk (first / last block of its bit map).
int run_chunks(spb_index* idx, DeviceCtx& c, const HostJob& job, const std::vector<spb::Chunk>& chunks, size_t c_begin, size_t c_end,
               uint64_t* total_sat, std::vector<EdgeRec>* edge_recs) {
    int rc = ensure_ctx(idx, c);
    if (rc) return rc;
    const int kSlots = idx->n_slots;
    const bool bits = job.width == 0;
    const uint64_t ptr_bytes = job.ptr32 ? 4 : 8;
    if (bits) { rc = c.h_edges.reserve((c_end - c_begin) * 64 + 64); if (rc) return rc; }
    for (int i = 0; i < kSlots; i++) c.slots[i].sat_pending = false;
    for (size_t ci = c_begin; ci < c_end; ci++) {
        const spb::Chunk& ch = chunks[ci];
        const uint64_t nr = ch.r1 - ch.r0, nb = ch.nb, b0 = ch.b0;
        Slot& s = c.slots[(ci - c_begin) % (size_t)kSlots];
        rc = ensure_slot_base(s);
        if (rc) return rc;
        SPB_CUDA(cudaStreamSynchronize(s.stream));  // previous use of this slot (incl. its pinned staging) is done
        if (s.sat_pending) { *total_sat += *s.h_sat; s.sat_pending = false; }
        // buffers (the stream is idle: growing is safe)
        const uint64_t bit_shift = bits ? (b0 & 255u) : 0;
        const uint64_t n_blk = bits ? (nb + bit_shift + 255) / 256 : 0;
        const size_t out_bytes = bits ? (size_t)n_blk * 32 : (size_t)nb * (size_t)job.width;
        if ((rc = s.d_reads.reserve((size_t)nb + 16, 512)) || (rc = s.d_offsets.reserve((size_t)nr + 1)) || (rc = s.h_offsets.reserve((size_t)nr + 1)) ||
            (rc = s.d_out.reserve(out_bytes + 64)) || (rc = s.d_stats.reserve(job.want_stats ? (size_t)nr : 0)))
            return rc;
        if (job.ms && (rc = s.d_pointers_raw.reserve(((size_t)nb * ptr_bytes + 7) / 8 + 8))) return rc;
        if (job.docs && (rc = s.d_docs.reserve((size_t)nb + 16))) return rc;
        for (uint64_t i = 0; i <= nr; i++) s.h_offsets.p[i] = job.offsets[ch.r0 + i] - b0;
        SPB_CUDA(cudaMemcpyAsync(s.d_offsets.p, s.h_offsets.p, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
        LaunchExtras ex;
        ex.bit_shift = bit_shift;
        // mixed read lengths: the kernel pulls the long reads first
        if ((rc = s.h_order.reserve((size_t)nr)) || (rc = s.d_order.reserve((size_t)nr))) return rc;
        if (spb::longest_first_order(job.offsets, ch, s.h_order.p)) {
            SPB_CUDA(cudaMemcpyAsync(s.d_order.p, s.h_order.p, nr * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
            ex.d_order = s.d_order.p;
        }
        if (nb && job.packed) {
            // packed base b0 + i = bit pair (b0 + i) of the caller's array: copy whole 32-bit words, the kernel shifts
            const uint64_t w0 = b0 / 16, shift = b0 % 16, n_words = (nb + shift + 15) / 16 + 1;
            if ((rc = s.d_packed.reserve((size_t)n_words))) return rc;
            SPB_CUDA(cudaMemcpyAsync(s.d_packed.p, job.packed + w0 * 4, n_words * 4, cudaMemcpyHostToDevice, s.stream));
            const uint64_t* e0 = std::lower_bound(job.exc_pos, job.exc_pos + job.n_exc, b0);
            const uint64_t* e1 = std::lower_bound(e0, job.exc_pos + job.n_exc, b0 + nb);
            const uint64_t ne = (uint64_t)(e1 - e0);
            if (ne) {
                if ((rc = s.d_exc_pos.reserve((size_t)ne)) || (rc = s.d_exc_val.reserve((size_t)ne))) return rc;
                SPB_CUDA(cudaMemcpyAsync(s.d_exc_pos.p, e0, ne * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
                SPB_CUDA(cudaMemcpyAsync(s.d_exc_val.p, job.exc_val + (e0 - job.exc_pos), ne, cudaMemcpyHostToDevice, s.stream));
            }
            int nl = 0;
            SPB_CUDA(spb::launch_unpack(s.d_packed.p, (uint32_t)shift, nb, s.d_reads.p, s.d_exc_pos.p, s.d_exc_val.p, ne, b0, c.num_sms, s.stream, &nl));
            g_launches.fetch_add((uint64_t)nl);
        } else if (nb)
            SPB_CUDA(cudaMemcpyAsync(s.d_reads.p, job.reads + b0, nb, cudaMemcpyHostToDevice, s.stream));
        if (job.count_sat) SPB_CUDA(cudaMemsetAsync(s.d_sat, 0, sizeof(unsigned long long), s.stream));
        if (bits && n_blk) SPB_CUDA(cudaMemsetAsync(s.d_out.p, 0, (size_t)n_blk * 32, s.stream));
        rc = launch_query(idx, c, job.ms, s.d_reads.p, s.d_offsets.p, nr, s.d_out.p, job.ms ? s.d_pointers_raw.p : nullptr, job.docs ? s.d_docs.p : nullptr,
                          job.want_stats ? s.d_stats.p : nullptr, job.opts, job.count_sat ? s.d_sat : nullptr, s.stream, ex);
        if (rc) return rc;
        if (bits) {
            // block g of the caller's map = bases [256 g, 256 g + 256).  The first and the last block of a sub-batch are shared
            // with its neighbours: they travel through the pinned edge buffer and are OR-ed in by query_host; the blocks in
            // between go straight to their place.
            const uint64_t g0 = (b0 - bit_shift) / 256;
            uint8_t* edge = c.h_edges.p + (ci - c_begin) * 64;
            (*edge_recs)[ci] = EdgeRec{g0, n_blk};
            if (n_blk) {
                SPB_CUDA(cudaMemcpyAsync(edge, s.d_out.p, 32, cudaMemcpyDeviceToHost, s.stream));
                if (n_blk > 1) SPB_CUDA(cudaMemcpyAsync(edge + 32, s.d_out.p + (n_blk - 1) * 32, 32, cudaMemcpyDeviceToHost, s.stream));
                if (n_blk > 2) SPB_CUDA(cudaMemcpyAsync((uint8_t*)job.lengths + (g0 + 1) * 32, s.d_out.p + 32, (n_blk - 2) * 32, cudaMemcpyDeviceToHost, s.stream));
            }
        } else if (nb)
            SPB_CUDA(cudaMemcpyAsync((uint8_t*)job.lengths + b0 * (uint64_t)job.width, s.d_out.p, nb * (uint64_t)job.width, cudaMemcpyDeviceToHost, s.stream));
        if (job.ms && nb) SPB_CUDA(cudaMemcpyAsync((uint8_t*)job.pointers + b0 * ptr_bytes, s.d_pointers_raw.p, nb * ptr_bytes, cudaMemcpyDeviceToHost, s.stream));
        if (job.docs && nb) SPB_CUDA(cudaMemcpyAsync(job.docs + b0, s.d_docs.p, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        if (job.want_stats) SPB_CUDA(cudaMemcpyAsync(job.stats + ch.r0, s.d_stats.p, nr * sizeof(spb_bin_stats), cudaMemcpyDeviceToHost, s.stream));
        if (job.count_sat) {
            SPB_CUDA(cudaMemcpyAsync(s.h_sat, s.d_sat, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
            s.sat_pending = true;
        }
    }
    for (int i = 0; i < kMaxSlots; i++) {
        Slot& s = c.slots[i];
        if (!s.stream) continue;
        SPB_CUDA(cudaStreamSynchronize(s.stream));
        if (s.sat_pending) { *total_sat += *s.h_sat; s.sat_pending = false; }
    }
    return SPB_OK;
}

int query_host(spb_index* idx, HostJob job, uint64_t n_reads) {
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    const spb_query_opts* opts = job.opts;
    if (opts && opts->saturated) *opts->saturated = 0;
    if (n_reads == 0) return SPB_OK;
    if ((!job.reads && !job.packed) || !job.offsets || !job.lengths || (job.ms && !job.pointers)) return fail(SPB_ERR_ARG, "null host buffer");
    if (job.packed && job.n_exc && (!job.exc_pos || !job.exc_val)) return fail(SPB_ERR_ARG, "null exception list");
    if (idx->devs.empty() || !idx->devs[0]->dev_blob) return fail(SPB_ERR_NO_DEVICE, "index is not resident on a GPU (call spb_index_to_device)");
    configure(idx);
    if (!width_ok(opts)) return fail(SPB_ERR_ARG, "length_bytes must be 0, 1, 2, 4 or SPB_LENGTH_BITS");
    job.width = out_width_of(opts);
    if (job.width == 0 && (job.ms || job.docs)) return fail(SPB_ERR_ARG, "SPB_LENGTH_BITS describes PML lengths only (no MS, no document numbers)");
    job.want_stats = job.stats && opts && opts->bin_width;
    job.ptr32 = job.ms && opts && (opts->flags & SPB_QUERY_POINTERS32);
    job.count_sat = (job.width == 1 || job.width == 2) || job.ptr32;
    DrainOnExit drain{idx};

    // sub-batches, then contiguous ranges of them per GPU, balanced by bases
    const size_t n_dev = idx->devs.size();
    std::vector<spb::Chunk> chunks;
    try {
        const uint64_t want = n_dev > 1 ? std::max<uint64_t>(1024, std::min<uint64_t>(idx->want_reads, n_reads / (n_dev * 4) + 1)) : idx->want_reads;
        chunks = spb::plan_chunks(job.offsets, 0, n_reads, want, idx->min_chunk_bases, idx->max_chunk_bases);
    } catch (std::length_error& e) {
        return fail(SPB_ERR_LIMIT, e.what());
    } catch (std::exception& e) {
        return fail(SPB_ERR_ARG, e.what());
    }
    std::vector<size_t> dev_cut(n_dev + 1, chunks.size());
    dev_cut[0] = 0;
    {
        const uint64_t base0 = job.offsets[0], total = job.offsets[n_reads] - base0;
        size_t ci = 0;
        for (size_t d = 1; d < n_dev; d++) {
            const uint64_t want = base0 + (uint64_t)((unsigned __int128)total * d / n_dev);
            while (ci < chunks.size() && chunks[ci].b0 + chunks[ci].nb / 2 < want) ci++;
            dev_cut[d] = ci;
        }
    }
    std::vector<EdgeRec> edge_recs(job.width == 0 ? chunks.size() : 0);
    std::vector<uint64_t> sat(n_dev, 0);
    std::vector<int> rcs(n_dev, SPB_OK);
    std::vector<std::string> errs(n_dev);
    if (n_dev == 1) {
        rcs[0] = run_chunks(idx, *idx->devs[0], job, chunks, 0, chunks.size(), &sat[0], &edge_recs);
        if (rcs[0]) return rcs[0];
    } else {
        // one host thread per GPU (the reference's `-t` region, compute_ms_pml.cpp:890-894, is the same shape: independent
        // workers over disjoint batches of reads sharing a read-only index)
        std::vector<std::thread> th;
        for (size_t d = 0; d < n_dev; d++)
            th.emplace_back([&, d] {
                rcs[d] = run_chunks(idx, *idx->devs[d], job, chunks, dev_cut[d], dev_cut[d + 1], &sat[d], &edge_recs);
                if (rcs[d]) errs[d] = g_error;
            });
        for (auto& t : th) t.join();
        for (size_t d = 0; d < n_dev; d++)
            if (rcs[d]) return fail(rcs[d], "GPU " + std::to_string(idx->devs[d]->device) + ": " + errs[d]);
    }
    if (job.width == 0) {
        // edge blocks, in sub-batch order: a block is assigned by the first sub-batch that touches it and OR-ed by the next
        uint64_t* out = static_cast<uint64_t*>(job.lengths);
        uint64_t last_blk = ~0ull;
        for (size_t d = 0; d < n_dev; d++)
            for (size_t ci = dev_cut[d]; ci < dev_cut[d + 1]; ci++) {
                const EdgeRec& e = edge_recs[ci];
                if (!e.n_blk) continue;
                const uint64_t* src = reinterpret_cast<const uint64_t*>(idx->devs[d]->h_edges.p + (ci - dev_cut[d]) * 64);
                for (int half = 0; half < (e.n_blk > 1 ? 2 : 1); half++) {
                    const uint64_t g = half ? e.first_blk + e.n_blk - 1 : e.first_blk;
                    for (int q = 0; q < 4; q++) {
                        if (g == last_blk) out[g * 4 + q] |= src[half * 4 + q];
                        else out[g * 4 + q] = src[half * 4 + q];
                    }
                    last_blk = g;
                }
            }
    }
    uint64_t total_sat = 0;
    for (uint64_t v : sat) total_sat += v;
    if (opts && opts->saturated) *opts->saturated = total_sat;
    // A clamped PML length loses nothing (lengths[j] != 0 iff the clamped value != 0, and the non-zero values count up from
    // the right: spb_pml_widen_lengths rebuilds them); clamped MS lengths and 32-bit pointers cannot be rebuilt.
    if (total_sat && job.ms)
        return fail(SPB_ERR_OVERFLOW, std::to_string(total_sat) + " value(s) exceed the narrow length / pointer output and were clamped; repeat with wider outputs");
    return SPB_OK;
}

// ---- PML query + on-device decimal formatting (first resident GPU) ----------------------------------------------------
struct DrainCtxOnExit {  // one GPU's streams only: other host threads may be driving the other GPUs of the handle
    DeviceCtx* c;
    ~DrainCtxOnExit() {
        cudaSetDevice(c->device);
        for (auto& s : c->slots)
            if (s.stream) cudaStreamSynchronize(s.stream);
    }
};
struct TextChunk { int slot; uint64_t nr, id_bytes; uint32_t length_bytes; };
enum { kTextLengths = 0, kTextPointers = 1, kTextDocs = 2 };

// PML or MS query + on-device decimal formatting of every requested output (K5) on one GPU of the handle.
int query_host_text(spb_index* idx, int gpu_slot, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads, const uint8_t* ids,
                    const uint64_t* id_offsets, spb_text_buf* out_lengths, spb_text_buf* out_pointers, spb_text_buf* out_docs, spb_bin_stats* stats,
                    const spb_query_opts* opts) {
    if (!idx || !out_lengths) return fail(SPB_ERR_ARG, "null index / lengths text");
    const bool ms = idx->mode == SPB_MODE_MS;
    spb_text_buf* outs[3] = {out_lengths, ms ? out_pointers : nullptr, out_docs};
    if (ms && !out_pointers) return fail(SPB_ERR_ARG, "an MS index needs a pointers text buffer");
    if (!ms && out_pointers) return fail(SPB_ERR_ARG, "pointers text requested from a PML index");
    for (auto* o : outs) if (o) o->bytes = 0;
    if (n_reads == 0) return SPB_OK;
    if (!reads || !offsets || !ids || !id_offsets) return fail(SPB_ERR_ARG, "null host buffer");
    for (auto* o : outs) if (o && !o->text) return fail(SPB_ERR_ARG, "null text buffer");
    if (gpu_slot < 0 || (size_t)gpu_slot >= idx->devs.size() || !idx->devs[(size_t)gpu_slot]->dev_blob)
        return fail(SPB_ERR_NO_DEVICE, "index is not resident on that GPU (call spb_index_to_device / spb_index_replicate)");
    DeviceCtx& c = *idx->devs[(size_t)gpu_slot];
    int rc = ensure_ctx(idx, c);
    if (rc) return rc;
    DrainCtxOnExit drain{&c};
    const bool want_stats = stats && opts && opts->bin_width;
    const bool want_docs = out_docs != nullptr;
    const int kSlots = idx->n_slots;
    spb_query_opts qo;
    std::memset(&qo, 0, sizeof(qo));
    if (opts) qo = *opts;
    qo.saturated = nullptr;
    qo.flags &= ~SPB_QUERY_POINTERS32;  // the text carries the reference's 64-bit pointers (wrap-arounds included, App. C-4)
    std::vector<spb::Chunk> chunks;
    try {
        chunks = spb::plan_chunks(offsets, 0, n_reads, idx->want_reads, idx->min_chunk_bases, idx->max_chunk_bases);
    } catch (std::length_error& e) {
        return fail(SPB_ERR_LIMIT, e.what());
    } catch (std::exception& e) {
        return fail(SPB_ERR_ARG, e.what());
    }
    uint64_t text_off[3] = {0, 0, 0};
    std::deque<TextChunk> inflight;

    auto stream_values = [&](Slot& s, int which, uint32_t length_bytes, const void** values, uint32_t* elem) {
        if (which == kTextLengths) { *values = s.d_out.p; *elem = length_bytes; }
        else if (which == kTextPointers) { *values = s.d_pointers_raw.p; *elem = 8; }
        else { *values = s.d_docs.p; *elem = 4; }
    };
    // stage 2 of the oldest chunk: its sizes are known -> format on the device, copy every text to its place
    auto finish = [&]() -> int {
        const TextChunk t = inflight.front();
        inflight.pop_front();
        Slot& s = c.slots[t.slot];
        SPB_CUDA(cudaStreamSynchronize(s.stream));
        bool too_small = false;
        for (int w = 0; w < 3; w++) {
            if (!outs[w]) continue;
            const uint64_t total = s.h_total[w];
            if (text_off[w] + total > outs[w]->capacity) too_small = true;
        }
        if (too_small) {
            for (int w = 0; w < 3; w++) if (outs[w]) outs[w]->bytes = text_off[w] + s.h_total[w];  // bytes needed so far
            return fail(SPB_ERR_LIMIT, "text buffer too small");
        }
        for (int w = 0; w < 3; w++) {
            if (!outs[w]) continue;
            const uint64_t total = s.h_total[w];
            int r2 = s.ts[w].d_text.reserve((size_t)(total + total / 4), 64);
            if (r2) return r2;
            spb::TextParams tp;
            stream_values(s, w, t.length_bytes, &tp.values, &tp.elem_bytes);
            tp.offsets = s.d_offsets.p; tp.ids = s.d_ids.p; tp.id_offsets = s.d_id_offsets.p;
            tp.n_reads = t.nr; tp.starts = s.ts[w].d_starts.p; tp.text = s.ts[w].d_text.p;
            SPB_CUDA(spb::launch_text_format(tp, c.num_sms, s.stream));
            g_launches.fetch_add(1);
            if (total) SPB_CUDA(cudaMemcpyAsync(outs[w]->text + text_off[w], s.ts[w].d_text.p, total, cudaMemcpyDeviceToHost, s.stream));
            text_off[w] += total;
        }
        return SPB_OK;
    };

    for (size_t ci = 0; ci < chunks.size(); ci++) {
        const spb::Chunk& ch = chunks[ci];
        const uint64_t nr = ch.r1 - ch.r0, nb = ch.nb, b0 = ch.b0;
        const int slot = (int)(ci % (size_t)kSlots);
        while (!inflight.empty() && (int)inflight.size() >= (kSlots + 1) / 2) { rc = finish(); if (rc) return rc; }  // also frees `slot`'s previous use
        Slot& s = c.slots[slot];
        if ((rc = ensure_slot_base(s))) return rc;
        SPB_CUDA(cudaStreamSynchronize(s.stream));
        if (id_offsets[ch.r1] < id_offsets[ch.r0]) return fail(SPB_ERR_ARG, "id_offsets are not non-decreasing");
        const uint64_t i0 = id_offsets[ch.r0], id_bytes = id_offsets[ch.r1] - i0;
        // a PML / MS length never exceeds the read length: uint16 is lossless for reads < 65536 bases (document numbers need uint32 lengths)
        qo.length_bytes = (ch.max_len < 65536 && !want_docs) ? 2 : 4;
        if ((rc = s.d_reads.reserve((size_t)nb + 16, 512)) || (rc = s.d_offsets.reserve((size_t)nr + 1)) || (rc = s.h_offsets.reserve((size_t)nr + 1)) ||
            (rc = s.d_out.reserve((size_t)nb * qo.length_bytes + 64)) || (rc = s.d_stats.reserve(want_stats ? (size_t)nr : 0)) ||
            (rc = s.d_id_offsets.reserve((size_t)nr + 1)) || (rc = s.h_id_offsets.reserve((size_t)nr + 1)) || (rc = s.d_ids.reserve((size_t)id_bytes + 64)))
            return rc;
        if (ms && (rc = s.d_pointers_raw.reserve((size_t)nb + 8))) return rc;
        if (want_docs && (rc = s.d_docs.reserve((size_t)nb + 16))) return rc;
        for (int w = 0; w < 3; w++)
            if (outs[w] && (rc = s.ts[w].d_starts.reserve((size_t)nr + 1))) return rc;
        if (!s.h_total) SPB_CUDA(cudaMallocHost(&s.h_total, 3 * sizeof(uint64_t)));
        for (uint64_t i = 0; i <= nr; i++) {
            s.h_offsets.p[i] = offsets[ch.r0 + i] - b0;
            const uint64_t io = id_offsets[ch.r0 + i];
            if (io < i0 || (i && io < id_offsets[ch.r0 + i - 1])) return fail(SPB_ERR_ARG, "id_offsets are not non-decreasing");
            s.h_id_offsets.p[i] = io - i0;
        }
        SPB_CUDA(cudaMemcpyAsync(s.d_offsets.p, s.h_offsets.p, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
        SPB_CUDA(cudaMemcpyAsync(s.d_id_offsets.p, s.h_id_offsets.p, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
        if (nb) SPB_CUDA(cudaMemcpyAsync(s.d_reads.p, reads + b0, nb, cudaMemcpyHostToDevice, s.stream));
        if (id_bytes) SPB_CUDA(cudaMemcpyAsync(s.d_ids.p, ids + i0, id_bytes, cudaMemcpyHostToDevice, s.stream));
        LaunchExtras ex;
        if ((rc = s.h_order.reserve((size_t)nr)) || (rc = s.d_order.reserve((size_t)nr))) return rc;
        if (spb::longest_first_order(offsets, ch, s.h_order.p)) {
            SPB_CUDA(cudaMemcpyAsync(s.d_order.p, s.h_order.p, nr * sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
            ex.d_order = s.d_order.p;
        }
        rc = launch_query(idx, c, ms, s.d_reads.p, s.d_offsets.p, nr, s.d_out.p, ms ? s.d_pointers_raw.p : nullptr, want_docs ? s.d_docs.p : nullptr,
                          want_stats ? s.d_stats.p : nullptr, &qo, nullptr, s.stream, ex);
        if (rc) return rc;
        for (int w = 0; w < 3; w++) {
            if (!outs[w]) continue;
            spb::TextParams tp;
            stream_values(s, w, qo.length_bytes, &tp.values, &tp.elem_bytes);
            tp.offsets = s.d_offsets.p; tp.ids = s.d_ids.p; tp.id_offsets = s.d_id_offsets.p;
            tp.n_reads = nr; tp.starts = s.ts[w].d_starts.p; tp.text = nullptr;
            SPB_CUDA(spb::launch_text_sizes(tp, c.num_sms, s.stream));
            g_launches.fetch_add(2);
            SPB_CUDA(cudaMemcpyAsync(s.h_total + w, s.ts[w].d_starts.p + nr, sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream));
        }
        if (want_stats) SPB_CUDA(cudaMemcpyAsync(stats + ch.r0, s.d_stats.p, nr * sizeof(spb_bin_stats), cudaMemcpyDeviceToHost, s.stream));
        inflight.push_back(TextChunk{slot, nr, id_bytes, qo.length_bytes});
    }
    while (!inflight.empty()) { rc = finish(); if (rc) return rc; }
    for (int i = 0; i < kMaxSlots; i++)
        if (c.slots[i].stream) SPB_CUDA(cudaStreamSynchronize(c.slots[i].stream));
    for (int w = 0; w < 3; w++) if (outs[w]) outs[w]->bytes = text_off[w];
    return SPB_OK;
}

template <typename F>
int guarded(F&& f) {
    try {
        return f();
    } catch (spb::IoError& e) {
        return fail(SPB_ERR_IO, e.what());
    } catch (spb::LimitError& e) {
        return fail(SPB_ERR_LIMIT, e.what());
    } catch (std::bad_alloc&) {
        return fail(SPB_ERR_LIMIT, "out of host memory");
    } catch (std::exception& e) {
        return fail(SPB_ERR_FORMAT, e.what());
    }
}

// ---- replication over several GPUs -------------------------------------------------------------------------------------
// One ncclBroadcast of the flat image from the first GPU to the others (NCCL over NVLink / NVSwitch; the library is loaded
// at run time so that nothing links against it), or, when NCCL is not available, a binary tree of peer-to-peer copies.
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    bool load() {
        if (std::getenv("SPUMONI_B200_NO_NCCL")) return false;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!lib) return false;
        CommInitAll = (int (*)(void**, int, const int*))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
        GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
        GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
        Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclBroadcast");
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Broadcast;
    }
};

int broadcast_image(std::vector<std::unique_ptr<DeviceCtx>>& devs, size_t bytes, std::string* how) {
    const int n = (int)devs.size();
    NcclApi nccl;
    if (nccl.load()) {
        std::vector<void*> comms((size_t)n, nullptr);
        std::vector<int> ids((size_t)n);
        for (int i = 0; i < n; i++) ids[(size_t)i] = devs[(size_t)i]->device;
        if (nccl.CommInitAll(comms.data(), n, ids.data()) == 0) {
            bool ok = nccl.GroupStart() == 0;
            for (int i = 0; ok && i < n; i++) {
                cudaSetDevice(ids[(size_t)i]);
                ok = nccl.Broadcast(devs[0]->dev_blob, (void*)devs[(size_t)i]->dev_blob, bytes, /*ncclChar*/ 0, 0, comms[(size_t)i], (cudaStream_t)0) == 0;
            }
            ok = (nccl.GroupEnd() == 0) && ok;
            for (int i = 0; i < n; i++) { cudaSetDevice(ids[(size_t)i]); if (cudaDeviceSynchronize() != cudaSuccess) ok = false; }
            for (int i = 0; i < n; i++) if (comms[(size_t)i]) nccl.CommDestroy(comms[(size_t)i]);
            if (ok) { *how = "ncclBroadcast"; return SPB_OK; }
            cudaGetLastError();
        }
    }
    // peer copies: after round t the first 2^t GPUs hold the image
    for (int have = 1; have < n; have *= 2) {
        for (int i = 0; i < have && have + i < n; i++)
            SPB_CUDA(cudaMemcpyPeerAsync((void*)devs[(size_t)(have + i)]->dev_blob, devs[(size_t)(have + i)]->device, devs[(size_t)i]->dev_blob, devs[(size_t)i]->device, bytes, 0));
        for (int i = 0; i < n; i++) { SPB_CUDA(cudaSetDevice(devs[(size_t)i]->device)); SPB_CUDA(cudaDeviceSynchronize()); }
    }
    *how = "cudaMemcpyPeer tree";
    return SPB_OK;
}

std::string g_replicate_how;

}  // namespace

extern "C" {

const char* spb_last_error(void) { return g_error.c_str(); }
int spb_abi_version(void) { return SPB_ABI_VERSION; }

int spb_index_open(const char* ref_file, int mode, spb_index** out) { return spb_index_open_ex(ref_file, mode, 0, out); }

int spb_index_open_ex(const char* ref_file, int mode, uint32_t flags, spb_index** out) {
    if (!ref_file || !out || (mode != SPB_MODE_PML && mode != SPB_MODE_MS) || (flags & ~SPB_OPEN_DOC)) return fail(SPB_ERR_ARG, "spb_index_open: bad argument");
    *out = nullptr;
    return guarded([&]() -> int {
        const bool ms = mode == SPB_MODE_MS;
        const std::string ref(ref_file);
        const bool timing = std::getenv("SPUMONI_B200_TIMING") != nullptr;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double t0 = now();
        spb::DecodedIndex d = spb::parse_index_file(ref + (ms ? ".thrbv.ms" : ".thrbv.spumoni"), ms);
        const double t1 = now();
        if (flags & SPB_OPEN_DOC) spb::load_doc_array(ref + ".doc", d);
        std::string text;
        if (ms) text = spb::load_reference_text(ref, d.n - 1);
        const double t2 = now();
        spb::balance_runs(d, spb::default_balance_params());
        std::unique_ptr<spb_index> idx(new spb_index());
        idx->mode = mode;
        const double t3 = now();
        idx->host_blob = spb::build_flat_image(d, ms, text);
        if (timing) std::fprintf(stderr, "[timing] index open: parse %.3f s | text/doc %.3f s | balance %.3f s (%llu runs -> %llu rows) | flatten %.3f s (%.1f MB)\n",
                                 t1 - t0, t2 - t1, t3 - t2, (unsigned long long)d.R_file, (unsigned long long)d.R, now() - t3, idx->host_blob.size() / 1e6);
        std::memcpy(&idx->hdr, idx->host_blob.data(), sizeof(SpbFlatHeader));
        *out = idx.release();
        return SPB_OK;
    });
}

int spb_index_verify(const char* ref_file, int mode) {
    if (!ref_file || (mode != SPB_MODE_PML && mode != SPB_MODE_MS)) return fail(SPB_ERR_ARG, "spb_index_verify: bad argument");
    return guarded([&]() -> int {
        setenv("SPUMONI_B200_VERIFY_INDEX", "1", 1);
        struct Unset { ~Unset() { unsetenv("SPUMONI_B200_VERIFY_INDEX"); } } unset;
        const bool ms = mode == SPB_MODE_MS;
        (void)spb::parse_index_file(std::string(ref_file) + (ms ? ".thrbv.ms" : ".thrbv.spumoni"), ms);
        return SPB_OK;
    });
}

int spb_index_stats(const spb_index* idx, uint64_t* n, uint64_t* r) {
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    if (n) *n = idx->hdr.n;
    if (r) *r = idx->hdr.R_file;  // the runs of the BWT as the reference counts them, not the rows of the flat image
    return SPB_OK;
}

int spb_index_flat(const spb_index* idx, const void** host_blob, uint64_t* bytes) {
    if (!idx || !host_blob || !bytes) return fail(SPB_ERR_ARG, "spb_index_flat: bad argument");
    if (idx->host_blob.empty()) return fail(SPB_ERR_ARG, "this index holds no host copy of the flat image");
    *host_blob = idx->host_blob.data();
    *bytes = idx->host_blob.size();
    return SPB_OK;
}

int spb_index_from_flat(const void* host_blob, uint64_t bytes, spb_index** out) {
    if (!host_blob || !out) return fail(SPB_ERR_ARG, "spb_index_from_flat: bad argument");
    *out = nullptr;
    return guarded([&]() -> int {
        spb::validate_flat_image(host_blob, bytes);
        std::unique_ptr<spb_index> idx(new spb_index());
        std::memcpy(&idx->hdr, host_blob, sizeof(SpbFlatHeader));
        idx->mode = (int)idx->hdr.mode;
        idx->host_blob.assign((const uint8_t*)host_blob, (const uint8_t*)host_blob + bytes);
        *out = idx.release();
        return SPB_OK;
    });
}

uint64_t spb_flat_header_bytes(void) { return sizeof(SpbFlatHeader); }

int spb_index_from_header(const void* header, uint64_t bytes, spb_index** out) {
    if (!header || !out) return fail(SPB_ERR_ARG, "spb_index_from_header: bad argument");
    *out = nullptr;
    if (bytes < sizeof(SpbFlatHeader)) return fail(SPB_ERR_ARG, "header shorter than spb_flat_header_bytes()");
    SpbFlatHeader h;
    std::memcpy(&h, header, sizeof(h));
    if (h.magic != SPB_FLAT_MAGIC || h.version != SPB_FLAT_VERSION) return fail(SPB_ERR_FORMAT, "not a spumoni_b200 flat image (magic/version)");
    spb_index* idx = new spb_index();
    idx->hdr = h;
    idx->mode = (int)h.mode;
    *out = idx;
    return SPB_OK;
}

int spb_index_to_device(spb_index* idx, int device) {
    const int devices[1] = {device};
    return spb_index_replicate(idx, devices, 1);
}

int spb_index_replicate(spb_index* idx, const int* devices, int n_devices) {
    if (!idx || !devices || n_devices < 1) return fail(SPB_ERR_ARG, "spb_index_replicate: bad argument");
    if (idx->host_blob.empty()) return fail(SPB_ERR_ARG, "no host image to upload");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(SPB_ERR_NO_DEVICE, "no CUDA device available; spumoni_b200 has no CPU fallback");
    for (int i = 0; i < n_devices; i++) {
        if (devices[i] < 0 || devices[i] >= count) return fail(SPB_ERR_ARG, "device ordinal out of range");
        for (int j = 0; j < i; j++) if (devices[j] == devices[i]) return fail(SPB_ERR_ARG, "device listed twice");
    }
    idx->devs.clear();  // a previous residency (streams, buffers, counters, image copies) goes away with its contexts
    const size_t bytes = idx->host_blob.size();
    std::vector<std::unique_ptr<DeviceCtx>> devs;
    for (int i = 0; i < n_devices; i++) {
        std::unique_ptr<DeviceCtx> c(new DeviceCtx());
        SPB_CUDA(cudaSetDevice(devices[i]));
        void* p = nullptr;
        SPB_CUDA(cudaMalloc(&p, bytes));
        c->device = devices[i];
        c->dev_blob = (const uint8_t*)p;
        c->own_dev = true;
        devs.push_back(std::move(c));
    }
    SPB_CUDA(cudaSetDevice(devices[0]));
    SPB_CUDA(cudaMemcpy((void*)devs[0]->dev_blob, idx->host_blob.data(), bytes, cudaMemcpyHostToDevice));
    if (n_devices > 1) {
        int rc = broadcast_image(devs, bytes, &g_replicate_how);
        if (rc) return rc;
    }
    idx->devs = std::move(devs);
    for (auto& c : idx->devs) { int rc = ensure_ctx(idx, *c); if (rc) return rc; }
    return SPB_OK;
}

int spb_index_device_count(const spb_index* idx) { return idx ? (int)idx->devs.size() : 0; }
const char* spb_replicate_method(void) { return g_replicate_how.c_str(); }

int spb_index_attach_device(spb_index* idx, int device, const void* dev_blob, uint64_t bytes) {
    if (!idx || !dev_blob) return fail(SPB_ERR_ARG, "spb_index_attach_device: bad argument");
    if (bytes != idx->hdr.total_bytes) return fail(SPB_ERR_ARG, "device image size does not match the index");
    if (((uintptr_t)dev_blob & 255u) != 0) return fail(SPB_ERR_ARG, "device image must be 256-byte aligned");
    SPB_CUDA(cudaSetDevice(device));
    SpbFlatHeader chk;
    SPB_CUDA(cudaMemcpy(&chk, dev_blob, sizeof(chk), cudaMemcpyDeviceToHost));
    if (chk.magic != SPB_FLAT_MAGIC || chk.total_bytes != idx->hdr.total_bytes || chk.n != idx->hdr.n || chk.R != idx->hdr.R)
        return fail(SPB_ERR_FORMAT, "device memory does not hold this index's flat image");
    idx->devs.clear();
    std::unique_ptr<DeviceCtx> c(new DeviceCtx());
    c->device = device;
    c->dev_blob = (const uint8_t*)dev_blob;
    c->own_dev = false;
    idx->devs.push_back(std::move(c));
    return ensure_device_state(idx);
}

void spb_index_close(spb_index* idx) {
    delete idx;  // contexts release their streams, buffers and image copies
}

int spb_pml_query_host(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                       void* lengths, spb_bin_stats* stats, const spb_query_opts* opts) {
    HostJob j;
    j.reads = reads; j.offsets = offsets; j.lengths = lengths; j.stats = stats; j.opts = opts;
    return query_host(idx, j, n_reads);
}
int spb_pml_query_host_packed(spb_index* idx, const uint8_t* packed, const uint64_t* exc_pos, const uint8_t* exc_val, uint64_t n_exc,
                              const uint64_t* offsets, uint64_t n_reads, void* lengths, spb_bin_stats* stats, const spb_query_opts* opts) {
    if (!packed && n_reads) return fail(SPB_ERR_ARG, "null packed reads");
    HostJob j;
    j.packed = packed; j.exc_pos = exc_pos; j.exc_val = exc_val; j.n_exc = n_exc;
    j.offsets = offsets; j.lengths = lengths; j.stats = stats; j.opts = opts;
    return query_host(idx, j, n_reads);
}
int spb_pml_query_host_doc(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                           uint32_t* lengths, uint32_t* doc_nums, spb_bin_stats* stats, const spb_query_opts* opts) {
    if (!doc_nums) return fail(SPB_ERR_ARG, "null doc_nums");
    HostJob j;
    j.reads = reads; j.offsets = offsets; j.lengths = lengths; j.docs = doc_nums; j.stats = stats; j.opts = opts;
    return query_host(idx, j, n_reads);
}
int spb_ms_query_host_doc(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                          uint32_t* lengths, uint64_t* pointers, uint32_t* doc_nums, spb_bin_stats* stats,
                          const spb_query_opts* opts) {
    if (!doc_nums) return fail(SPB_ERR_ARG, "null doc_nums");
    HostJob j;
    j.ms = true; j.reads = reads; j.offsets = offsets; j.lengths = lengths; j.pointers = pointers; j.docs = doc_nums; j.stats = stats; j.opts = opts;
    return query_host(idx, j, n_reads);
}
uint32_t spb_index_num_docs(const spb_index* idx) { return idx ? idx->hdr.num_docs : 0; }
int spb_pml_query_host_text(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                            const uint8_t* ids, const uint64_t* id_offsets, char* text, uint64_t text_capacity,
                            uint64_t* text_bytes, spb_bin_stats* stats, const spb_query_opts* opts) {
    return spb_pml_query_host_text_on(idx, 0, reads, offsets, n_reads, ids, id_offsets, text, text_capacity, text_bytes, stats, opts);
}
int spb_pml_query_host_text_on(spb_index* idx, int gpu_slot, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                               const uint8_t* ids, const uint64_t* id_offsets, char* text, uint64_t text_capacity,
                               uint64_t* text_bytes, spb_bin_stats* stats, const spb_query_opts* opts) {
    if (!text_bytes) return fail(SPB_ERR_ARG, "null text_bytes");
    *text_bytes = 0;
    if (idx && idx->mode != SPB_MODE_PML) return fail(SPB_ERR_ARG, "spb_pml_query_host_text on an MS index");
    spb_text_buf t{text, text_capacity, 0};
    const int rc = query_host_text(idx, gpu_slot, reads, offsets, n_reads, ids, id_offsets, &t, nullptr, nullptr, stats, opts);
    *text_bytes = t.bytes;
    return rc;
}
int spb_query_host_text_on(spb_index* idx, int gpu_slot, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                           const uint8_t* ids, const uint64_t* id_offsets, spb_text_buf* lengths_text, spb_text_buf* pointers_text,
                           spb_text_buf* docs_text, spb_bin_stats* stats, const spb_query_opts* opts) {
    return query_host_text(idx, gpu_slot, reads, offsets, n_reads, ids, id_offsets, lengths_text, pointers_text, docs_text, stats, opts);
}
int spb_ms_query_host(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                      void* lengths, void* pointers, spb_bin_stats* stats, const spb_query_opts* opts) {
    HostJob j;
    j.ms = true; j.reads = reads; j.offsets = offsets; j.lengths = lengths; j.pointers = pointers; j.stats = stats; j.opts = opts;
    return query_host(idx, j, n_reads);
}
int spb_pml_query_device(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                         void* d_lengths, spb_bin_stats* d_stats, const spb_query_opts* opts, void* cuda_stream) {
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    return launch_query(idx, *primary(idx), false, d_reads, d_offsets, n_reads, d_lengths, nullptr, nullptr, d_stats, opts,
                        opts ? (unsigned long long*)opts->saturated : nullptr, (cudaStream_t)cuda_stream, LaunchExtras{});
}
int spb_pml_query_device_doc(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                             uint32_t* d_lengths, uint32_t* d_doc_nums, spb_bin_stats* d_stats,
                             const spb_query_opts* opts, void* cuda_stream) {
    if (!idx || !d_doc_nums) return fail(SPB_ERR_ARG, "null index / doc_nums");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    return launch_query(idx, *primary(idx), false, d_reads, d_offsets, n_reads, d_lengths, nullptr, d_doc_nums, d_stats, opts, nullptr, (cudaStream_t)cuda_stream, LaunchExtras{});
}
int spb_ms_query_device_doc(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                            uint32_t* d_lengths, uint64_t* d_pointers, uint32_t* d_doc_nums, spb_bin_stats* d_stats,
                            const spb_query_opts* opts, void* cuda_stream) {
    if (!idx || !d_doc_nums) return fail(SPB_ERR_ARG, "null index / doc_nums");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    return launch_query(idx, *primary(idx), true, d_reads, d_offsets, n_reads, d_lengths, d_pointers, d_doc_nums, d_stats, opts, nullptr, (cudaStream_t)cuda_stream, LaunchExtras{});
}
int spb_ms_query_device(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                        void* d_lengths, void* d_pointers, spb_bin_stats* d_stats,
                        const spb_query_opts* opts, void* cuda_stream) {
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    return launch_query(idx, *primary(idx), true, d_reads, d_offsets, n_reads, d_lengths, (uint64_t*)d_pointers, nullptr, d_stats, opts,
                        opts ? (unsigned long long*)opts->saturated : nullptr, (cudaStream_t)cuda_stream, LaunchExtras{});
}

int spb_nulldb_load(const char* path, spb_nulldb* out) {
    if (!path || !out) return fail(SPB_ERR_ARG, "spb_nulldb_load: bad argument");
    return guarded([&]() -> int {
        spb::NullDb db = spb::load_null_db(path);
        out->num_values = db.num_values;
        out->ks_stat_threshold = db.ks_stat_threshold;
        out->mean_null_stat = db.mean_null_stat;
        out->percentile_value = db.percentile_value;
        return SPB_OK;
    });
}

int spb_nulldb_generate_host(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                             uint32_t* null_stats, spb_nulldb* summary) {
    if (!idx || !offsets || !summary || (n_reads && (!reads || !null_stats))) return fail(SPB_ERR_ARG, "spb_nulldb_generate_host: bad argument");
    return guarded([&]() -> int {
        // null read = the read reversed (compute_ms_pml.cpp:1425-1426 / :1463-1464); ::toupper commutes with the reversal
        const uint64_t total = offsets[n_reads];
        std::vector<uint8_t> rev((size_t)total);
        for (uint64_t r = 0; r < n_reads; r++) {
            const uint64_t a = offsets[r], b = offsets[r + 1];
            if (b < a || b > total) return fail(SPB_ERR_ARG, "offsets are not non-decreasing");
            for (uint64_t i = a; i < b; i++) rev[(size_t)i] = reads[a + (b - 1 - i)];
        }
        spb_query_opts o;
        std::memset(&o, 0, sizeof(o));
        o.flags = SPB_QUERY_TOUPPER;
        HostJob j;
        j.reads = rev.data(); j.offsets = offsets; j.lengths = null_stats; j.opts = &o;
        std::vector<uint64_t> ptrs;
        if (idx->mode == SPB_MODE_MS) {
            ptrs.resize((size_t)total);
            j.ms = true; j.pointers = ptrs.data();
        }
        const int rc = query_host(idx, j, n_reads);
        if (rc) return rc;
        const spb::NullDb db = spb::summarize_null_stats(null_stats, total);
        summary->num_values = db.num_values;
        summary->ks_stat_threshold = 0.0;
        summary->mean_null_stat = db.mean_null_stat;
        summary->percentile_value = db.percentile_value;
        return SPB_OK;
    });
}

int spb_nulldb_save(const char* path, const spb_nulldb* summary, const uint32_t* null_stats) {
    if (!path || !summary || (summary->num_values && !null_stats)) return fail(SPB_ERR_ARG, "spb_nulldb_save: bad argument");
    return guarded([&]() -> int {
        spb::NullDb db;
        db.num_values = summary->num_values;
        db.ks_stat_threshold = summary->ks_stat_threshold;
        db.mean_null_stat = summary->mean_null_stat;
        db.percentile_value = summary->percentile_value;
        spb::save_null_db(path, db, null_stats);
        return SPB_OK;
    });
}

uint64_t spb_max_value_threshold(double percentile_value, int is_pml, int use_dna_letters, int use_promotions) {
    uint64_t thr = (uint64_t)(percentile_value > 3.0 ? percentile_value : 3.0);
    if (use_dna_letters) thr++;
    else if (is_pml && !use_promotions) thr += 4;
    return thr;
}

int spb_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(SPB_ERR_ARG, "spb_host_alloc: null");
    SPB_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));  // portable: pinned for every GPU of a replicated index
    return SPB_OK;
}
void spb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int spb_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
}
uint64_t spb_kernel_launches(void) { return g_launches.load(); }

// ---- index construction on the GPU -------------------------------------------------------------
int spb_build_parts_device(const uint8_t* text, uint64_t n, int device, spb_build_parts* out) {
    if (!text || !out) return fail(SPB_ERR_ARG, "spb_build_parts_device: bad argument");
    std::memset(out, 0, sizeof(*out));
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(SPB_ERR_NO_DEVICE, "no CUDA device available; spumoni_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail(SPB_ERR_ARG, "device ordinal out of range");
    for (uint64_t i = 0; i + 1 < n; i++) if (text[i] <= 1) return fail(SPB_ERR_ARG, "the text contains a byte <= 0x01 before its terminator");
    return guarded([&]() -> int {
        spb::BuildParts P;
        std::string err;
        if (!spb::build_rlbwt_device(text, n, device, P, err, std::getenv("SPUMONI_B200_TIMING") != nullptr))
            return fail(err.find("length") != std::string::npos ? SPB_ERR_LIMIT : SPB_ERR_CUDA, err);
        const size_t r = P.heads.size();
        out->n = P.n; out->r = r;
        out->heads = (uint8_t*)std::malloc(r ? r : 1);
        uint64_t* block = (uint64_t*)std::malloc((r ? r : 1) * 4 * sizeof(uint64_t));
        if (!out->heads || !block) { std::free(out->heads); std::free(block); std::memset(out, 0, sizeof(*out)); return fail(SPB_ERR_LIMIT, "out of host memory"); }
        out->lens = block; out->thr = block + r; out->ssa = block + 2 * r; out->esa = block + 3 * r;
        std::memcpy(out->heads, P.heads.data(), r);
        std::memcpy(out->lens, P.lens.data(), r * 8);
        std::memcpy(out->thr, P.thr.data(), r * 8);
        std::memcpy(out->ssa, P.ssa.data(), r * 8);
        std::memcpy(out->esa, P.esa.data(), r * 8);
        return SPB_OK;
    });
}
void spb_build_parts_free(spb_build_parts* p) {
    if (!p) return;
    std::free(p->heads);
    std::free(p->lens);  // lens / thr / ssa / esa share one allocation
    std::memset(p, 0, sizeof(*p));
}
int spb_build_parts_write(const char* ref_file, const spb_build_parts* p) {
    if (!ref_file || !p || (p->r && (!p->heads || !p->lens))) return fail(SPB_ERR_ARG, "spb_build_parts_write: bad argument");
    const std::string ref(ref_file);
    auto put5 = [](std::FILE* f, uint64_t v) { std::fwrite(&v, 1, 5, f); };
    std::FILE* h = std::fopen((ref + ".bwt.heads").c_str(), "wb");
    std::FILE* l = std::fopen((ref + ".bwt.len").c_str(), "wb");
    std::FILE* t = std::fopen((ref + ".thr_pos").c_str(), "wb");
    std::FILE* s = std::fopen((ref + ".ssa").c_str(), "wb");
    std::FILE* e = std::fopen((ref + ".esa").c_str(), "wb");
    const bool ok = h && l && t && s && e;
    if (ok) {
        uint64_t pos = 0;
        std::fwrite(p->heads, 1, p->r, h);
        for (uint64_t k = 0; k < p->r; k++) {
            put5(l, p->lens[k]);
            put5(t, p->thr[k]);
            put5(s, pos); put5(s, p->ssa[k]);
            pos += p->lens[k];
            put5(e, pos - 1); put5(e, p->esa[k]);
        }
    }
    bool closed = true;
    for (std::FILE* f : {h, l, t, s, e}) if (f && std::fclose(f) != 0) closed = false;
    if (!ok || !closed) return fail(SPB_ERR_IO, "cannot write the build intermediates next to " + ref);
    return SPB_OK;
}

// test hooks for the host-side planning (no GPU needed)
int spb_plan_chunks(const uint64_t* offsets, uint64_t n_reads, uint64_t want_reads, uint64_t min_bases, uint64_t max_bases,
                    uint64_t* chunk_first_read, uint64_t capacity, uint64_t* n_chunks) {
    if (!offsets || !n_chunks) return fail(SPB_ERR_ARG, "spb_plan_chunks: bad argument");
    try {
        const std::vector<spb::Chunk> c = spb::plan_chunks(offsets, 0, n_reads, want_reads, min_bases, max_bases);
        *n_chunks = c.size();
        if (c.size() + 1 > capacity || !chunk_first_read) return c.empty() ? SPB_OK : fail(SPB_ERR_LIMIT, "capacity");
        for (size_t i = 0; i < c.size(); i++) chunk_first_read[i] = c[i].r0;
        chunk_first_read[c.size()] = n_reads;
        return SPB_OK;
    } catch (std::length_error& e) {
        return fail(SPB_ERR_LIMIT, e.what());
    } catch (std::exception& e) {
        return fail(SPB_ERR_ARG, e.what());
    }
}
int spb_plan_order(const uint64_t* offsets, uint64_t n_reads, uint32_t* order, int* sorted) {
    if (!offsets || !order || !sorted) return fail(SPB_ERR_ARG, "spb_plan_order: bad argument");
    spb::Chunk c;
    c.r0 = 0; c.r1 = n_reads; c.b0 = offsets[0]; c.nb = offsets[n_reads] - offsets[0];
    c.min_len = ~0ull; c.max_len = 0;
    for (uint64_t i = 0; i < n_reads; i++) { const uint64_t l = offsets[i + 1] - offsets[i]; c.min_len = std::min(c.min_len, l); c.max_len = std::max(c.max_len, l); }
    *sorted = spb::longest_first_order(offsets, c, order) ? 1 : 0;
    return SPB_OK;
}

}  // extern "C"
