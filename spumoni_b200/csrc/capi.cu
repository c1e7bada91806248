// C ABI of libspumoni_b200.so (declared in include/spumoni_b200.h).
//
// Host plumbing only: index lifetime, device residency, batch pipelining (H2D / kernel / D2H on
// rotating streams) and error translation.  All per-base work happens in kernels.cu.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <deque>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/spumoni_b200.h"
#include "flat_index.h"
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

// Host entry points: the batch is cut into sub-batches that rotate over `n_slots` streams.  One read is one thread,
// so a sub-batch of 1 kb reads only fills a fraction of the ~150 000 resident lanes: several sub-batches must be in
// flight (copies of some overlapping the kernels of others) to keep both the SMs and the PCIe link busy.
constexpr int kMaxSlots = 16;
constexpr int kDefaultSlots = 8;                       // SPUMONI_B200_SLOTS
constexpr uint64_t kDefaultChunkBases = 32ull << 20;   // SPUMONI_B200_CHUNK_MB (in Mi bases)
constexpr int kCounters = 64;     // pairs of work counters, rotated per launch
constexpr int kDefaultLanes = 4;  // cooperative kernel (indexes without fast symbols)

struct Slot {
    cudaStream_t stream = nullptr;
    uint8_t* d_reads = nullptr;
    uint64_t* d_offsets = nullptr;
    uint32_t* d_lengths = nullptr;
    uint64_t* d_pointers = nullptr;
    uint32_t* d_docs = nullptr;
    spb_bin_stats* d_stats = nullptr;
    uint64_t* h_offsets = nullptr;  // pinned staging for chunk-local offsets
    unsigned long long* d_sat = nullptr;  // clamped-value counter of the launch in flight
    unsigned long long* h_sat = nullptr;  // pinned copy
    bool sat_pending = false;
    uint64_t cap_bases = 0, cap_reads = 0;
    bool cap_ms = false, cap_doc = false;
    // on-device text formatting (spb_pml_query_host_text)
    uint8_t* d_ids = nullptr;
    uint64_t *d_id_offsets = nullptr, *d_starts = nullptr, *h_id_offsets = nullptr, *h_total = nullptr;
    char* d_text = nullptr;
    uint64_t cap_ids = 0, cap_text = 0, cap_text_reads = 0;
};

}  // namespace

struct spb_index {
    int mode = 0;
    SpbFlatHeader hdr;
    spb::FlatBlob host_blob;
    int device = -1;
    const uint8_t* dev_blob = nullptr;
    bool own_dev = false;
    int num_sms = 0;
    unsigned long long* d_counters = nullptr;
    uint32_t counter_next = 0;
    uint32_t row_ld = 1;  // move rows allocate in L1 (the neighbour rows of a jump / normalisation hit there); SPUMONI_B200_ROW_LD
    Slot slots[kMaxSlots];
    int n_slots = kDefaultSlots;
    uint64_t chunk_bases = kDefaultChunkBases;
};

namespace {

int ensure_device_state(spb_index* idx) {
    if (idx->device < 0 || !idx->dev_blob) return fail(SPB_ERR_NO_DEVICE, "index is not resident on a GPU (call spb_index_to_device)");
    SPB_CUDA(cudaSetDevice(idx->device));
    if (!idx->d_counters) {
        SPB_CUDA(cudaMalloc(&idx->d_counters, kCounters * sizeof(unsigned long long)));
        SPB_CUDA(cudaDeviceGetAttribute(&idx->num_sms, cudaDevAttrMultiProcessorCount, idx->device));
        // Every backward-search step is one random 16/32-byte gather; the default L2 fill granularity
        // drags whole 128-byte lines out of HBM for each of them (ncu: 128 B of DRAM traffic per
        // step).  Ask for sector-sized fills.  SPUMONI_B200_L2_FETCH=0 keeps the device default.
        size_t gran = 32;
        if (const char* e = std::getenv("SPUMONI_B200_L2_FETCH")) gran = (size_t)std::atoi(e);
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);  // a hint: failure is not an error
        cudaGetLastError();
        if (const char* e = std::getenv("SPUMONI_B200_ROW_LD")) idx->row_ld = (uint32_t)std::atoi(e);  // development knob
        if (const char* e = std::getenv("SPUMONI_B200_SLOTS")) { int v = std::atoi(e); if (v >= 1 && v <= kMaxSlots) idx->n_slots = v; }
        if (const char* e = std::getenv("SPUMONI_B200_CHUNK_MB")) { long v = std::atol(e); if (v >= 1 && v <= 4096) idx->chunk_bases = (uint64_t)v << 20; }
    }
    return SPB_OK;
}

void free_slot(Slot& s) {
    if (s.d_reads) cudaFree(s.d_reads);
    if (s.d_offsets) cudaFree(s.d_offsets);
    if (s.d_lengths) cudaFree(s.d_lengths);
    if (s.d_pointers) cudaFree(s.d_pointers);
    if (s.d_docs) cudaFree(s.d_docs);
    if (s.d_stats) cudaFree(s.d_stats);
    if (s.h_offsets) cudaFreeHost(s.h_offsets);
    if (s.d_sat) cudaFree(s.d_sat);
    if (s.h_sat) cudaFreeHost(s.h_sat);
    if (s.d_ids) cudaFree(s.d_ids);
    if (s.d_id_offsets) cudaFree(s.d_id_offsets);
    if (s.d_starts) cudaFree(s.d_starts);
    if (s.d_text) cudaFree(s.d_text);
    if (s.h_id_offsets) cudaFreeHost(s.h_id_offsets);
    if (s.h_total) cudaFreeHost(s.h_total);
    s.d_ids = nullptr; s.d_id_offsets = nullptr; s.d_starts = nullptr; s.d_text = nullptr; s.h_id_offsets = nullptr; s.h_total = nullptr;
    s.cap_ids = s.cap_text = s.cap_text_reads = 0;
    s.d_sat = nullptr; s.h_sat = nullptr; s.sat_pending = false;
    s.d_reads = nullptr; s.d_offsets = nullptr; s.d_lengths = nullptr; s.d_pointers = nullptr; s.d_docs = nullptr; s.d_stats = nullptr;
    s.h_offsets = nullptr;
    s.cap_bases = s.cap_reads = 0;
}

int ensure_slot(Slot& s, uint64_t bases, uint64_t reads, bool ms, bool doc) {
    if (!s.stream) SPB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (bases <= s.cap_bases && reads <= s.cap_reads && (!ms || s.cap_ms) && (!doc || s.cap_doc)) return SPB_OK;
    ms = ms || s.cap_ms;
    doc = doc || s.cap_doc;
    SPB_CUDA(cudaStreamSynchronize(s.stream));
    const uint64_t nb = bases > s.cap_bases ? bases : s.cap_bases, nr = reads > s.cap_reads ? reads : s.cap_reads;
    cudaStream_t keep = s.stream;
    free_slot(s);
    s.stream = keep;
    SPB_CUDA(cudaMalloc(&s.d_reads, nb + 512));
    SPB_CUDA(cudaMalloc(&s.d_offsets, (nr + 1) * sizeof(uint64_t)));
    SPB_CUDA(cudaMalloc(&s.d_lengths, (nb + 64) * sizeof(uint32_t)));
    SPB_CUDA(cudaMalloc(&s.d_stats, (nr + 1) * sizeof(spb_bin_stats)));
    if (ms) SPB_CUDA(cudaMalloc(&s.d_pointers, (nb + 64) * sizeof(uint64_t)));
    if (doc) SPB_CUDA(cudaMalloc(&s.d_docs, (nb + 64) * sizeof(uint32_t)));
    SPB_CUDA(cudaMallocHost(&s.h_offsets, (nr + 1) * sizeof(uint64_t)));
    SPB_CUDA(cudaMalloc(&s.d_sat, sizeof(unsigned long long)));
    SPB_CUDA(cudaMallocHost(&s.h_sat, sizeof(unsigned long long)));
    s.cap_bases = nb; s.cap_reads = nr; s.cap_ms = ms; s.cap_doc = doc;
    return SPB_OK;
}

// The host entry points hand caller memory to asynchronous copies: whatever way they return (errors included), no
// copy may still be in flight afterwards.
struct DrainOnExit {
    spb_index* idx;
    ~DrainOnExit() {
        for (int i = 0; i < kMaxSlots; i++)
            if (idx->slots[i].stream) cudaStreamSynchronize(idx->slots[i].stream);
    }
};

int length_bytes_of(const spb_query_opts* opts) {
    const uint32_t w = opts ? opts->length_bytes : 0;
    return w == 0 ? 4 : (int)w;
}

int launch_query(spb_index* idx, bool ms, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                 void* d_lengths, uint64_t* d_pointers, uint32_t* d_docs, spb_bin_stats* d_stats, const spb_query_opts* opts,
                 unsigned long long* d_saturated, cudaStream_t stream) {
    if (n_reads == 0) return SPB_OK;
    if ((ms ? 1 : 0) != idx->mode) return fail(SPB_ERR_ARG, ms ? "MS query on a PML index" : "PML query on an MS index");
    if (!d_reads || !d_offsets || !d_lengths || (ms && !d_pointers)) return fail(SPB_ERR_ARG, "null device buffer");
    if (((uintptr_t)d_reads & 255u) != 0) return fail(SPB_ERR_ARG, "d_reads must be 256-byte aligned");
    if (((uintptr_t)d_lengths & 15u) != 0 || ((uintptr_t)d_pointers & 15u) != 0) return fail(SPB_ERR_ARG, "d_lengths / d_pointers must be 16-byte aligned");
    const SpbFlatHeader& h = idx->hdr;
    const bool has_fast = h.num_fast > 0 && (h.off_mrows != 0 || h.off_mrows8 != 0) && (!ms || h.off_msamp != 0 || h.mrows8_stride == 16);
    int lanes = opts && opts->lanes_per_read ? (int)opts->lanes_per_read : (has_fast ? 1 : kDefaultLanes);
    if (lanes != 1 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) return fail(SPB_ERR_ARG, "lanes_per_read must be 1, 4, 8, 16 or 32");
    if (lanes == 1 && !has_fast) return fail(SPB_ERR_ARG, "lanes_per_read = 1 needs an index with fast symbols (frequent run heads < 0x80)");
    const int width = length_bytes_of(opts);
    if (width != 1 && width != 2 && width != 4) return fail(SPB_ERR_ARG, "length_bytes must be 0, 1, 2 or 4");
    if (width != 4 && lanes != 1) return fail(SPB_ERR_ARG, "narrow length output needs lanes_per_read = 1");
    if (d_docs) {
        if (!h.num_docs || !h.off_cdoc) return fail(SPB_ERR_ARG, "document numbers requested but the index was opened without SPB_OPEN_DOC");
        if (lanes != 1 || !h.off_mdoc) return fail(SPB_ERR_ARG, "document numbers need the thread-per-read kernels (lanes_per_read = 1, an index with fast symbols)");
        if (width != 4) return fail(SPB_ERR_ARG, "document numbers need 4-byte lengths");
        if (((uintptr_t)d_docs & 15u) != 0) return fail(SPB_ERR_ARG, "d_doc_nums must be 16-byte aligned");
    }
    spb::WalkParams p;
    std::memset(&p, 0, sizeof(p));
    p.runs = reinterpret_cast<const uint4*>(idx->dev_blob + h.off_runs);
    p.blocks = idx->dev_blob + h.off_blocks;
    p.crec = reinterpret_cast<const uint4*>(idx->dev_blob + h.off_crec);
    p.csamp = ms ? reinterpret_cast<const uint4*>(idx->dev_blob + h.off_csamp) : nullptr;
    p.text = ms ? idx->dev_blob + h.off_text : nullptr;
    p.chartab = reinterpret_cast<const SpbCharTab*>(idx->dev_blob + offsetof(SpbFlatHeader, chartab));
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
    p.rows = (has_fast && h.off_mrows) ? reinterpret_cast<const uint4*>(idx->dev_blob + h.off_mrows) : nullptr;
    p.rows8 = (has_fast && h.off_mrows8) ? reinterpret_cast<const uint64_t*>(idx->dev_blob + h.off_mrows8) : nullptr;
    p.msamp = (has_fast && ms && h.off_msamp) ? reinterpret_cast<const uint4*>(idx->dev_blob + h.off_msamp) : nullptr;
    p.rows8_stride = h.mrows8_stride;
    for (int i = 0; i < 4; i++) { p.mv_end_k[i] = h.mv_end_k[i]; p.mv_end_off[i] = h.mv_end_off[i]; p.mv_end_sl[i] = h.mv_end_sl[i]; }
    p.out_width = (uint32_t)width;
    p.row_ld = idx->row_ld;
    p.saturated = d_saturated;
    p.docs = d_docs;
    if (d_docs) {
        p.cdoc = reinterpret_cast<const uint2*>(idx->dev_blob + h.off_cdoc);
        p.mdoc = reinterpret_cast<const uint2*>(idx->dev_blob + h.off_mdoc);
        p.doc_init = h.doc_init; p.doc_absent_ms = h.doc_absent_ms;
        for (int i = 0; i < 4; i++) p.mv_end_doc[i] = h.mv_end_doc[i];
    }
    p.work_counter = idx->d_counters + 2 * (idx->counter_next++ % (kCounters / 2));
    SPB_CUDA(cudaMemsetAsync(p.work_counter, 0, 2 * sizeof(unsigned long long), stream));
    int launched = 0;
    SPB_CUDA(spb::launch_walk(p, ms, lanes, idx->num_sms, stream, &launched));
    g_launches.fetch_add((uint64_t)launched);
    return SPB_OK;
}

int query_host(spb_index* idx, bool ms, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
               void* lengths, void* pointers, uint32_t* docs, spb_bin_stats* stats, const spb_query_opts* opts) {
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    if (opts && opts->saturated) *opts->saturated = 0;
    if (n_reads == 0) return SPB_OK;
    if (!reads || !offsets || !lengths || (ms && !pointers)) return fail(SPB_ERR_ARG, "null host buffer");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    DrainOnExit drain{idx};
    const bool want_stats = stats && opts && opts->bin_width;
    const int width = length_bytes_of(opts);
    if (width != 1 && width != 2 && width != 4) return fail(SPB_ERR_ARG, "length_bytes must be 0, 1, 2 or 4");
    const bool ptr32 = ms && opts && (opts->flags & SPB_QUERY_POINTERS32);
    const bool count_sat = width != 4 || ptr32;
    const uint64_t ptr_bytes = ptr32 ? 4 : 8;
    uint64_t total_sat = 0;
    uint64_t r0 = 0;
    int chunk = 0;
    const int kSlots = idx->n_slots;
    const uint64_t kChunkBases = idx->chunk_bases;
    while (r0 < n_reads) {
        // sub-batch [r0, r1): at least one read, up to kChunkBases bases
        uint64_t r1 = r0 + 1;
        while (r1 < n_reads && offsets[r1 + 1] - offsets[r0] <= kChunkBases) r1++;
        const uint64_t b0 = offsets[r0], nb = offsets[r1] - b0, nr = r1 - r0;
        if (offsets[r1] < b0) return fail(SPB_ERR_ARG, "offsets are not non-decreasing");
        Slot& s = idx->slots[chunk % kSlots];
        rc = ensure_slot(s, nb > kChunkBases ? nb : kChunkBases, nr > (kChunkBases / 64) ? nr : (kChunkBases / 64), ms, docs != nullptr);
        if (rc) return rc;
        SPB_CUDA(cudaStreamSynchronize(s.stream));  // previous use of this slot (incl. its pinned staging) is done
        if (s.sat_pending) { total_sat += *s.h_sat; s.sat_pending = false; }
        for (uint64_t i = 0; i <= nr; i++) {
            const uint64_t o = offsets[r0 + i];
            if (o < b0 || (i && o < offsets[r0 + i - 1])) return fail(SPB_ERR_ARG, "offsets are not non-decreasing");
            if (i && o - offsets[r0 + i - 1] > 0xFFFFFFFFull) return fail(SPB_ERR_LIMIT, "read longer than 2^32-1 bases");
            s.h_offsets[i] = o - b0;
        }
        SPB_CUDA(cudaMemcpyAsync(s.d_offsets, s.h_offsets, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
        if (nb) SPB_CUDA(cudaMemcpyAsync(s.d_reads, reads + b0, nb, cudaMemcpyHostToDevice, s.stream));
        if (count_sat) SPB_CUDA(cudaMemsetAsync(s.d_sat, 0, sizeof(unsigned long long), s.stream));
        rc = launch_query(idx, ms, s.d_reads, s.d_offsets, nr, s.d_lengths, s.d_pointers, docs ? s.d_docs : nullptr, s.d_stats, opts, count_sat ? s.d_sat : nullptr, s.stream);
        if (rc) return rc;
        if (nb) SPB_CUDA(cudaMemcpyAsync((uint8_t*)lengths + b0 * (uint64_t)width, s.d_lengths, nb * (uint64_t)width, cudaMemcpyDeviceToHost, s.stream));
        if (ms && nb) SPB_CUDA(cudaMemcpyAsync((uint8_t*)pointers + b0 * ptr_bytes, s.d_pointers, nb * ptr_bytes, cudaMemcpyDeviceToHost, s.stream));
        if (docs && nb) SPB_CUDA(cudaMemcpyAsync(docs + b0, s.d_docs, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        if (want_stats) SPB_CUDA(cudaMemcpyAsync(stats + r0, s.d_stats, nr * sizeof(spb_bin_stats), cudaMemcpyDeviceToHost, s.stream));
        if (count_sat) {
            SPB_CUDA(cudaMemcpyAsync(s.h_sat, s.d_sat, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
            s.sat_pending = true;
        }
        r0 = r1;
        chunk++;
    }
    for (int i = 0; i < kMaxSlots; i++) {
        Slot& s = idx->slots[i];
        if (!s.stream) continue;
        SPB_CUDA(cudaStreamSynchronize(s.stream));
        if (s.sat_pending) { total_sat += *s.h_sat; s.sat_pending = false; }
    }
    if (opts && opts->saturated) *opts->saturated = total_sat;
    if (total_sat) return fail(SPB_ERR_OVERFLOW, std::to_string(total_sat) + " value(s) exceed the narrow length / pointer output and were clamped; repeat with wider outputs");
    return SPB_OK;
}

// buffers of the text path; grown on demand (the slot's stream is idle when this is called)
int ensure_text_slot(Slot& s, uint64_t reads, uint64_t id_bytes, uint64_t text_bytes) {
    if (reads > s.cap_text_reads) {
        if (s.d_id_offsets) cudaFree(s.d_id_offsets);
        if (s.d_starts) cudaFree(s.d_starts);
        if (s.h_id_offsets) cudaFreeHost(s.h_id_offsets);
        s.d_id_offsets = s.d_starts = s.h_id_offsets = nullptr;
        const uint64_t nr = reads + reads / 4;
        SPB_CUDA(cudaMalloc(&s.d_id_offsets, (nr + 1) * sizeof(uint64_t)));
        SPB_CUDA(cudaMalloc(&s.d_starts, (nr + 1) * sizeof(uint64_t)));
        SPB_CUDA(cudaMallocHost(&s.h_id_offsets, (nr + 1) * sizeof(uint64_t)));
        s.cap_text_reads = nr;
    }
    if (!s.h_total) SPB_CUDA(cudaMallocHost(&s.h_total, sizeof(uint64_t)));
    if (id_bytes > s.cap_ids) {
        if (s.d_ids) cudaFree(s.d_ids);
        s.d_ids = nullptr;
        SPB_CUDA(cudaMalloc(&s.d_ids, id_bytes + id_bytes / 4 + 64));
        s.cap_ids = id_bytes + id_bytes / 4;
    }
    if (text_bytes > s.cap_text) {
        if (s.d_text) cudaFree(s.d_text);
        s.d_text = nullptr;
        SPB_CUDA(cudaMalloc(&s.d_text, text_bytes + text_bytes / 4 + 64));
        s.cap_text = text_bytes + text_bytes / 4;
    }
    return SPB_OK;
}

struct TextChunk { int slot; uint64_t nr, id_bytes; uint32_t elem_bytes; };

int query_host_text(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads, const uint8_t* ids,
                    const uint64_t* id_offsets, char* text, uint64_t text_capacity, uint64_t* text_bytes, spb_bin_stats* stats,
                    const spb_query_opts* opts) {
    if (!idx || !text_bytes) return fail(SPB_ERR_ARG, "null index / text_bytes");
    *text_bytes = 0;
    if (n_reads == 0) return SPB_OK;
    if (!reads || !offsets || !ids || !id_offsets || !text) return fail(SPB_ERR_ARG, "null host buffer");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    DrainOnExit drain{idx};
    const bool want_stats = stats && opts && opts->bin_width;
    const int kSlots = idx->n_slots;
    const uint64_t kChunkBases = idx->chunk_bases;
    spb_query_opts qo;
    std::memset(&qo, 0, sizeof(qo));
    if (opts) qo = *opts;
    qo.saturated = nullptr;
    uint64_t text_off = 0;
    std::deque<TextChunk> inflight;

    // stage 2 of the oldest chunk: its sizes are known -> format on the device, copy the text to its place
    auto finish = [&]() -> int {
        const TextChunk c = inflight.front();
        inflight.pop_front();
        Slot& s = idx->slots[c.slot];
        SPB_CUDA(cudaStreamSynchronize(s.stream));
        const uint64_t total = *s.h_total;
        if (text_off + total > text_capacity) { *text_bytes = text_off + total; return fail(SPB_ERR_LIMIT, "text buffer too small"); }
        int r2 = ensure_text_slot(s, c.nr, c.id_bytes, total);
        if (r2) return r2;
        spb::TextParams tp;
        tp.values = s.d_lengths; tp.offsets = s.d_offsets; tp.ids = s.d_ids; tp.id_offsets = s.d_id_offsets;
        tp.n_reads = c.nr; tp.starts = s.d_starts; tp.text = s.d_text; tp.elem_bytes = c.elem_bytes;
        SPB_CUDA(spb::launch_text_format(tp, idx->num_sms, s.stream));
        g_launches.fetch_add(1);
        if (total) SPB_CUDA(cudaMemcpyAsync(text + text_off, s.d_text, total, cudaMemcpyDeviceToHost, s.stream));
        text_off += total;
        return SPB_OK;
    };

    uint64_t r0 = 0;
    int chunk = 0;
    while (r0 < n_reads) {
        uint64_t r1 = r0 + 1;
        while (r1 < n_reads && offsets[r1 + 1] - offsets[r0] <= kChunkBases) r1++;
        const uint64_t b0 = offsets[r0], nb = offsets[r1] - b0, nr = r1 - r0;
        if (offsets[r1] < b0) return fail(SPB_ERR_ARG, "offsets are not non-decreasing");
        const int slot = chunk % kSlots;
        while (!inflight.empty() && (int)inflight.size() >= (kSlots + 1) / 2) { rc = finish(); if (rc) return rc; }  // also frees `slot`'s previous use
        Slot& s = idx->slots[slot];
        rc = ensure_slot(s, nb > kChunkBases ? nb : kChunkBases, nr > (kChunkBases / 64) ? nr : (kChunkBases / 64), false, false);
        if (rc) return rc;
        SPB_CUDA(cudaStreamSynchronize(s.stream));
        const uint64_t i0 = id_offsets[r0], id_bytes = id_offsets[r1] - i0;
        rc = ensure_text_slot(s, nr, id_bytes, 0);
        if (rc) return rc;
        uint64_t max_len = 0;
        for (uint64_t i = 0; i <= nr; i++) {
            const uint64_t o = offsets[r0 + i];
            if (o < b0 || (i && o < offsets[r0 + i - 1])) return fail(SPB_ERR_ARG, "offsets are not non-decreasing");
            if (i && o - offsets[r0 + i - 1] > max_len) max_len = o - offsets[r0 + i - 1];
            s.h_offsets[i] = o - b0;
            const uint64_t io = id_offsets[r0 + i];
            if (io < i0 || (i && io < id_offsets[r0 + i - 1])) return fail(SPB_ERR_ARG, "id_offsets are not non-decreasing");
            s.h_id_offsets[i] = io - i0;
        }
        if (max_len > 0xFFFFFFFFull) return fail(SPB_ERR_LIMIT, "read longer than 2^32-1 bases");
        qo.length_bytes = max_len < 65536 ? 2 : 4;  // a PML never exceeds the read length: uint16 is lossless here
        SPB_CUDA(cudaMemcpyAsync(s.d_offsets, s.h_offsets, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
        SPB_CUDA(cudaMemcpyAsync(s.d_id_offsets, s.h_id_offsets, (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
        if (nb) SPB_CUDA(cudaMemcpyAsync(s.d_reads, reads + b0, nb, cudaMemcpyHostToDevice, s.stream));
        if (id_bytes) SPB_CUDA(cudaMemcpyAsync(s.d_ids, ids + i0, id_bytes, cudaMemcpyHostToDevice, s.stream));
        rc = launch_query(idx, false, s.d_reads, s.d_offsets, nr, s.d_lengths, nullptr, nullptr, s.d_stats, &qo, nullptr, s.stream);
        if (rc) return rc;
        spb::TextParams tp;
        tp.values = s.d_lengths; tp.offsets = s.d_offsets; tp.ids = s.d_ids; tp.id_offsets = s.d_id_offsets;
        tp.n_reads = nr; tp.starts = s.d_starts; tp.text = nullptr; tp.elem_bytes = qo.length_bytes;
        SPB_CUDA(spb::launch_text_sizes(tp, idx->num_sms, s.stream));
        g_launches.fetch_add(2);
        SPB_CUDA(cudaMemcpyAsync(s.h_total, s.d_starts + nr, sizeof(uint64_t), cudaMemcpyDeviceToHost, s.stream));
        if (want_stats) SPB_CUDA(cudaMemcpyAsync(stats + r0, s.d_stats, nr * sizeof(spb_bin_stats), cudaMemcpyDeviceToHost, s.stream));
        inflight.push_back(TextChunk{slot, nr, id_bytes, qo.length_bytes});
        r0 = r1;
        chunk++;
    }
    while (!inflight.empty()) { rc = finish(); if (rc) return rc; }
    for (int i = 0; i < kMaxSlots; i++)
        if (idx->slots[i].stream) SPB_CUDA(cudaStreamSynchronize(idx->slots[i].stream));
    *text_bytes = text_off;
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
        if (ms) text = spb::load_reference_text(ref);
        std::unique_ptr<spb_index> idx(new spb_index());
        idx->mode = mode;
        const double t2 = now();
        idx->host_blob = spb::build_flat_image(d, ms, text);
        if (timing) std::fprintf(stderr, "[timing] index open: parse %.3f s | text/doc %.3f s | flatten %.3f s (%.1f MB)\n", t1 - t0, t2 - t1, now() - t2, idx->host_blob.size() / 1e6);
        std::memcpy(&idx->hdr, idx->host_blob.data(), sizeof(SpbFlatHeader));
        *out = idx.release();
        return SPB_OK;
    });
}

int spb_index_stats(const spb_index* idx, uint64_t* n, uint64_t* r) {
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    if (n) *n = idx->hdr.n;
    if (r) *r = idx->hdr.R;
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
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    if (idx->host_blob.empty()) return fail(SPB_ERR_ARG, "no host image to upload");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(SPB_ERR_NO_DEVICE, "no CUDA device available; spumoni_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail(SPB_ERR_ARG, "device ordinal out of range");
    SPB_CUDA(cudaSetDevice(device));
    if (idx->own_dev && idx->dev_blob) cudaFree((void*)idx->dev_blob);
    void* p = nullptr;
    SPB_CUDA(cudaMalloc(&p, idx->host_blob.size()));
    SPB_CUDA(cudaMemcpy(p, idx->host_blob.data(), idx->host_blob.size(), cudaMemcpyHostToDevice));
    idx->dev_blob = (const uint8_t*)p;
    idx->own_dev = true;
    idx->device = device;
    return ensure_device_state(idx);
}

int spb_index_attach_device(spb_index* idx, int device, const void* dev_blob, uint64_t bytes) {
    if (!idx || !dev_blob) return fail(SPB_ERR_ARG, "spb_index_attach_device: bad argument");
    if (bytes != idx->hdr.total_bytes) return fail(SPB_ERR_ARG, "device image size does not match the index");
    if (((uintptr_t)dev_blob & 255u) != 0) return fail(SPB_ERR_ARG, "device image must be 256-byte aligned");
    SPB_CUDA(cudaSetDevice(device));
    SpbFlatHeader chk;
    SPB_CUDA(cudaMemcpy(&chk, dev_blob, sizeof(chk), cudaMemcpyDeviceToHost));
    if (chk.magic != SPB_FLAT_MAGIC || chk.total_bytes != idx->hdr.total_bytes || chk.n != idx->hdr.n || chk.R != idx->hdr.R)
        return fail(SPB_ERR_FORMAT, "device memory does not hold this index's flat image");
    if (idx->own_dev && idx->dev_blob) cudaFree((void*)idx->dev_blob);
    idx->dev_blob = (const uint8_t*)dev_blob;
    idx->own_dev = false;
    idx->device = device;
    return ensure_device_state(idx);
}

void spb_index_close(spb_index* idx) {
    if (!idx) return;
    if (idx->device >= 0) {
        cudaSetDevice(idx->device);
        for (int i = 0; i < kMaxSlots; i++) {
            if (idx->slots[i].stream) cudaStreamSynchronize(idx->slots[i].stream);
            free_slot(idx->slots[i]);
            if (idx->slots[i].stream) cudaStreamDestroy(idx->slots[i].stream);
        }
        if (idx->d_counters) cudaFree(idx->d_counters);
        if (idx->own_dev && idx->dev_blob) cudaFree((void*)idx->dev_blob);
    }
    delete idx;
}

int spb_pml_query_host(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                       void* lengths, spb_bin_stats* stats, const spb_query_opts* opts) {
    return query_host(idx, false, reads, offsets, n_reads, lengths, nullptr, nullptr, stats, opts);
}
int spb_pml_query_host_doc(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                           uint32_t* lengths, uint32_t* doc_nums, spb_bin_stats* stats, const spb_query_opts* opts) {
    if (!doc_nums) return fail(SPB_ERR_ARG, "null doc_nums");
    return query_host(idx, false, reads, offsets, n_reads, lengths, nullptr, doc_nums, stats, opts);
}
int spb_ms_query_host_doc(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                          uint32_t* lengths, uint64_t* pointers, uint32_t* doc_nums, spb_bin_stats* stats,
                          const spb_query_opts* opts) {
    if (!doc_nums) return fail(SPB_ERR_ARG, "null doc_nums");
    return query_host(idx, true, reads, offsets, n_reads, lengths, pointers, doc_nums, stats, opts);
}
uint32_t spb_index_num_docs(const spb_index* idx) { return idx ? idx->hdr.num_docs : 0; }
int spb_pml_query_host_text(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                            const uint8_t* ids, const uint64_t* id_offsets, char* text, uint64_t text_capacity,
                            uint64_t* text_bytes, spb_bin_stats* stats, const spb_query_opts* opts) {
    if (idx && idx->mode != SPB_MODE_PML) return fail(SPB_ERR_ARG, "spb_pml_query_host_text on an MS index");
    return query_host_text(idx, reads, offsets, n_reads, ids, id_offsets, text, text_capacity, text_bytes, stats, opts);
}
int spb_ms_query_host(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                      void* lengths, void* pointers, spb_bin_stats* stats, const spb_query_opts* opts) {
    return query_host(idx, true, reads, offsets, n_reads, lengths, pointers, nullptr, stats, opts);
}
int spb_pml_query_device(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                         void* d_lengths, spb_bin_stats* d_stats, const spb_query_opts* opts, void* cuda_stream) {
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    return launch_query(idx, false, d_reads, d_offsets, n_reads, d_lengths, nullptr, nullptr, d_stats, opts,
                        opts ? (unsigned long long*)opts->saturated : nullptr, (cudaStream_t)cuda_stream);
}
int spb_pml_query_device_doc(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                             uint32_t* d_lengths, uint32_t* d_doc_nums, spb_bin_stats* d_stats,
                             const spb_query_opts* opts, void* cuda_stream) {
    if (!idx || !d_doc_nums) return fail(SPB_ERR_ARG, "null index / doc_nums");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    return launch_query(idx, false, d_reads, d_offsets, n_reads, d_lengths, nullptr, d_doc_nums, d_stats, opts, nullptr, (cudaStream_t)cuda_stream);
}
int spb_ms_query_device_doc(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                            uint32_t* d_lengths, uint64_t* d_pointers, uint32_t* d_doc_nums, spb_bin_stats* d_stats,
                            const spb_query_opts* opts, void* cuda_stream) {
    if (!idx || !d_doc_nums) return fail(SPB_ERR_ARG, "null index / doc_nums");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    return launch_query(idx, true, d_reads, d_offsets, n_reads, d_lengths, d_pointers, d_doc_nums, d_stats, opts, nullptr, (cudaStream_t)cuda_stream);
}
int spb_ms_query_device(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                        void* d_lengths, void* d_pointers, spb_bin_stats* d_stats,
                        const spb_query_opts* opts, void* cuda_stream) {
    if (!idx) return fail(SPB_ERR_ARG, "null index");
    int rc = ensure_device_state(idx);
    if (rc) return rc;
    return launch_query(idx, true, d_reads, d_offsets, n_reads, d_lengths, (uint64_t*)d_pointers, nullptr, d_stats, opts,
                        opts ? (unsigned long long*)opts->saturated : nullptr, (cudaStream_t)cuda_stream);
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
        int rc;
        if (idx->mode == SPB_MODE_MS) {
            std::vector<uint64_t> ptrs((size_t)total);
            rc = query_host(idx, true, rev.data(), offsets, n_reads, null_stats, ptrs.data(), nullptr, nullptr, &o);
        } else
            rc = query_host(idx, false, rev.data(), offsets, n_reads, null_stats, nullptr, nullptr, nullptr, &o);
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
    SPB_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
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

}  // extern "C"
