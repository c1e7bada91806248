/*
 * spumoni_b200.h -- C ABI of the B200-native MS/PML hot path.
 *
 * This is the drop-in boundary for SPUMONI's per-read query path.  The reference exposes no
 * FFI; the in-process seam these entry points replace is
 *     pml_t::matching_statistics(const char*, size_t, std::vector<size_t>&)
 *         /root/reference/src/compute_ms_pml.cpp:730-732
 *     ms_t::matching_statistics(const char*, size_t, std::vector<size_t>&, std::vector<size_t>&)
 *         /root/reference/src/compute_ms_pml.cpp:795-812
 * which the reference calls once per read from N OpenMP threads (:936-938, :1124-1126).  A GPU
 * cannot be fed per read, so the ABI is batch-granular: reads are passed concatenated, one byte
 * per base, with an offsets array (CSR style).  Each entry point cites the reference interface it
 * replaces.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - every function returns 0 on success and a non-zero spb_status on failure; the library
 *     never calls exit() (the reference's FATAL_* macros do, spumoni_main.hpp:28-33) -- the
 *     message is available from spb_last_error() (thread-local);
 *   - the caller owns all input/output buffers; index objects are owned by the library;
 *   - "ref_file" is the path the reference passes around internally: "<prefix>.fa"
 *     (spumoni.cpp:744-747); the index is "<ref_file>.thrbv.spumoni" / "<ref_file>.thrbv.ms"
 *     (compute_ms_pml.cpp:215-217, :544-546), exactly as written by `spumoni build`;
 *   - the CUDA path is the only path: without a usable GPU the query calls fail loudly;
 *   - device-wide settings this library touches: cudaLimitPersistingL2CacheSize is raised to its maximum on every GPU an index
 *     becomes resident on (PML walks over row tables that nearly fit the L2 pin them there: +4-6 %); persisting lines are
 *     reset when the index is closed, SPUMONI_B200_L2_PERSIST=0 leaves the limit alone.  cudaLimitMaxL2FetchGranularity is
 *     only touched when SPUMONI_B200_L2_FETCH is set;
 *   - an spb_index serialises nothing: use one host thread per handle (the reference shares one index between its
 *     OpenMP threads because every call handles one read; here one call handles a whole batch, keeps several
 *     sub-batches in flight on its own streams and, for an index replicated on several GPUs, drives one host thread
 *     per GPU itself).  Different handles are independent.
 */
#ifndef SPUMONI_B200_H
#define SPUMONI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPB_ABI_VERSION 4
#define SPB_QUERY_TOUPPER 1u
#define SPB_QUERY_POINTERS32 2u /* MS queries: `pointers` is a uint32_t array.  Lossless whenever every pointer is below 2^32 - 1
                                   (text positions of an index with n < 2^32, reads without characters absent from the index --
                                   the reference's size_t wrap-around after such a character, SURVEY App. C-4, does not fit).
                                   A value that does not fit is clamped, counted in `saturated`, and the *_host call returns
                                   SPB_ERR_OVERFLOW: discard the outputs and repeat with 64-bit pointers. */
#define SPB_LENGTH_BITS 0x80u /* spb_query_opts.length_bytes: PML lengths as one bit per base (see spb_pml_expand_bits) */
#define SPB_OPEN_DOC 1u /* spb_index_open_ex: also load <ref_file>.doc (the reference's use_doc / -d) */

typedef enum {
    SPB_OK = 0,
    SPB_ERR_IO = 1,        /* file missing / unreadable */
    SPB_ERR_FORMAT = 2,    /* on-disk structure failed validation */
    SPB_ERR_ARG = 3,       /* bad argument */
    SPB_ERR_CUDA = 4,      /* CUDA runtime error (message has the cudaError string) */
    SPB_ERR_NO_DEVICE = 5, /* index not resident on a GPU / no GPU present */
    SPB_ERR_LIMIT = 6,     /* size outside supported range (n >= 2^40-2, r >= 2^32-64, read >= 2^32) */
    SPB_ERR_OVERFLOW = 7   /* MS queries only: a narrow output (1/2-byte lengths, 32-bit pointers) clamped at least one value; the
                              outputs are otherwise complete: repeat the call with a wider length_bytes / without
                              SPB_QUERY_POINTERS32.  (Clamped PML lengths lose nothing and are not an error, see length_bytes.) */
} spb_status;

typedef enum { SPB_MODE_PML = 0, SPB_MODE_MS = 1 } spb_mode;

typedef struct spb_index spb_index; /* opaque: pml_t / ms_t equivalent */

/* Per-read result of the bin-max classifier (compute_ms_pml.cpp:969-995 / :1150-1176).
 * nbins = above + below; status FOUND iff above/(above+below) > 0.5; avg = sum_max / nbins. */
typedef struct {
    uint64_t sum_max; /* sum of per-bin maxima              (sum_max_bin_values) */
    uint32_t above;   /* bins with max >= max_value_thr     (bins_above)         */
    uint32_t below;   /* bins with max <  max_value_thr     (bins_below)         */
} spb_bin_stats;

typedef struct {
    uint32_t bin_width;      /* -w; 0 disables classification (stats may then be NULL) */
    uint32_t lanes_per_read; /* 0 = library default; 1 = thread-per-read kernels over the move rows (needs an index whose
                                frequent run heads are < 0x80, e.g. DNA), 4/8/16/32 = cooperative generic kernel */
    uint64_t max_value_thr;  /* see spb_max_value_threshold */
    uint32_t flags;          /* SPB_QUERY_TOUPPER: apply ::toupper to read bytes on the device (compute_ms_pml.cpp:916-917);
                                SPB_QUERY_POINTERS32: see above */
    uint32_t length_bytes;   /* element size of the `lengths` output: 0 or 4 = uint32_t, 2 = uint16_t, 1 = uint8_t.  A PML/MS
                                length never exceeds the read length, so 2 is lossless for reads < 65536 bases.  Values that do
                                not fit are clamped to the maximum and counted (lanes_per_read 1 only).  For PML a clamped
                                array is still lossless -- lengths[j] = (position j matched) ? lengths[j + 1] + 1 : 0
                                (compute_ms_pml.cpp:249-281), and clamping keeps zero / non-zero -- so the call succeeds and
                                spb_pml_widen_lengths rebuilds the exact values where `saturated` says some were clamped.
                                SPB_LENGTH_BITS (PML, lanes_per_read 1): `lengths` is a bit map of spb_bits_bytes(total bases)
                                bytes, bit i (bit i % 64 of 64-bit word i / 64) set iff the PML at base i is non-zero:
                                1/8 byte per base over PCIe for any read length; spb_pml_expand_bits turns it into lengths. */
    uint64_t* saturated;     /* optional.  *_host calls: host pointer, receives the number of clamped values (MS calls then
                                return SPB_ERR_OVERFLOW).  *_device calls: device pointer, incremented by the kernels. */
} spb_query_opts;

/* EmpNullDatabase (emp_null_database.hpp:16-33); values are not kept, only the scalars */
typedef struct {
    uint64_t num_values;
    double ks_stat_threshold;
    double mean_null_stat;
    double percentile_value;
} spb_nulldb;

const char* spb_last_error(void);
int spb_abi_version(void);

/* ---- index lifetime ------------------------------------------------------------------ */

/* pml_t::pml_t(filename, ...) compute_ms_pml.cpp:700-721 -> pml_pointers::load :222-229
 * ms_t::ms_t(filename, ...)   compute_ms_pml.cpp:755-786 -> ms_pointers::load  :551-562
 * Parses the reference's on-disk index (host only, no GPU needed) and flattens it into the
 * device layout.  MS mode takes the text from <ref_file> itself instead of <ref_file>.slp: the sequence lines of
 * the FASTA written by `spumoni build`, or the file verbatim for an index built with `build -g` (the BWT length
 * decides) -- see DESIGN.md. */
int spb_index_open(const char* ref_file, int mode, spb_index** out);
/* Same with flags.  SPB_OPEN_DOC: pml_t / ms_t constructed with use_doc = true (compute_ms_pml.cpp:712-720, :777-785)
 * -> DocumentArray::load of <ref_file>.doc (doc_array.cpp:197-201); enables the *_doc queries below. */
int spb_index_open_ex(const char* ref_file, int mode, uint32_t flags, spb_index** out);

/* Deep check of an index file (no GPU needed): parses it like spb_index_open and additionally RE-DERIVES every redundant
 * acceleration structure stored in it -- sdsl's select_support_mcl (two per Elias-Fano vector and per wavelet tree) and
 * rank_support_v -- from the decoded bits and compares them with the file byte for byte.  SPB_ERR_FORMAT names the field
 * that disagrees.  This is the self-test to run on the first index written by a real `spumoni build`: the byte layout of
 * the third-party containers (sdsl-lite, r-index) is reconstructed from their published formats, not from vendored sources
 * (docs/FORMAT_MAP.md lists every field with the upstream function that writes it). */
int spb_index_verify(const char* ref_file, int mode);

/* get_bwt_stats() compute_ms_pml.cpp:171-173 / :484-486 */
int spb_index_stats(const spb_index* idx, uint64_t* n, uint64_t* r);

/* The flat image: one contiguous, position-independent blob (header + tables).  Multi-GPU
 * loaders broadcast these bytes (NCCL) instead of re-parsing the file on every rank. */
int spb_index_flat(const spb_index* idx, const void** host_blob, uint64_t* bytes);
int spb_index_from_flat(const void* host_blob, uint64_t bytes, spb_index** out);
/* Header-only variant for ranks that receive the image directly into device memory: the first
 * spb_flat_header_bytes() bytes of the image describe it completely; pair with
 * spb_index_attach_device. */
uint64_t spb_flat_header_bytes(void);
int spb_index_from_header(const void* header, uint64_t bytes, spb_index** out);

/* Make the index resident on `device`: library-owned allocation + H2D copy ... */
int spb_index_to_device(spb_index* idx, int device);
/* ... or on several GPUs at once: one H2D copy to devices[0], then ONE broadcast of the flat image to the others over
 * NVLink (ncclBroadcast when libnccl can be loaded, peer-to-peer copies otherwise; spb_replicate_method() says which).
 * Afterwards every *_host query of this handle cuts its reads into contiguous ranges balanced by bases, one range per
 * GPU, driven by one host thread per GPU, no collective in the steady state -- the shape of the reference's `-t` region
 * (`#pragma omp parallel` over batches of reads on a shared read-only index, compute_ms_pml.cpp:890-894, :1080-1084).
 * The *_device entry points and spb_pml_query_host_text keep using devices[0].  Replaces any previous residency. */
int spb_index_replicate(spb_index* idx, const int* devices, int n_devices);
int spb_index_device_count(const spb_index* idx); /* GPUs the index is resident on */
const char* spb_replicate_method(void);           /* "ncclBroadcast" | "cudaMemcpyPeer tree" | "" (last replicate of this process) */
/* ... or adopt caller-owned device memory that already holds the flat image (e.g. the
 * receive buffer of an ncclBroadcast).  The memory must stay valid until spb_index_close. */
int spb_index_attach_device(spb_index* idx, int device, const void* dev_blob, uint64_t bytes);

void spb_index_close(spb_index* idx);

/* ---- queries ------------------------------------------------------------------------- */
/* Reads: `reads` holds n_reads reads back to back, one byte per base (already upper-cased or
 * not: the kernels apply ::toupper as compute_ms_pml.cpp:916-917 does when `toupper` != 0 in
 * the *_host calls; device calls take bytes verbatim); read i is reads[offsets[i] .. offsets[i+1]).
 * Outputs are indexed like the input: lengths[offsets[i] + j] is position j of read i. */

/* pml_t::matching_statistics compute_ms_pml.cpp:730-732 -> pml_pointers::_query :237-286,
 * plus (optional) the bin-max classifier :969-995.  Host buffers; H2D + kernel + D2H inside;
 * returns when the outputs are complete.  Pinned host memory (spb_host_alloc) makes the
 * copies asynchronous and overlapped. */
int spb_pml_query_host(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                       void* lengths /* uint32_t[], or uint16_t[] / uint8_t[] per opts->length_bytes */, spb_bin_stats* stats,
                       const spb_query_opts* opts);

/* Same query, reads passed as 2 bits per base (spb_pack_reads): a quarter of the host-to-device bytes.  packed / exc_*
 * are what spb_pack_reads produced for the concatenated reads; offsets / outputs as in spb_pml_query_host. */
int spb_pml_query_host_packed(spb_index* idx, const uint8_t* packed, const uint64_t* exc_pos, const uint8_t* exc_val, uint64_t n_exc,
                              const uint64_t* offsets, uint64_t n_reads, void* lengths, spb_bin_stats* stats, const spb_query_opts* opts);

/* ms_t::matching_statistics compute_ms_pml.cpp:795-812 -> ms_pointers::_query :570-623 */
int spb_ms_query_host(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                      void* lengths, void* pointers /* uint64_t[], or uint32_t[] with SPB_QUERY_POINTERS32 */, spb_bin_stats* stats,
                      const spb_query_opts* opts);

/* Same, on buffers already resident on the index's device; asynchronous on `cuda_stream`
 * (a cudaStream_t passed as void*; NULL = default stream).  d_reads must be 256-byte aligned
 * and readable up to the next multiple of 256 bytes past offsets[n_reads]; d_lengths must be 16-byte aligned
 * (SPB_LENGTH_BITS: 32-byte aligned, spb_bits_bytes(total bases) bytes, ZEROED by the caller before the call).
 * At most 1024 launches of one handle may be in flight at a time (work counters are recycled after that many). */
int spb_pml_query_device(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                         void* d_lengths, spb_bin_stats* d_stats, const spb_query_opts* opts, void* cuda_stream);
int spb_ms_query_device(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                        void* d_lengths, void* d_pointers, spb_bin_stats* d_stats,
                        const spb_query_opts* opts, void* cuda_stream);

/* PML query + on-device decimal formatting (K5): instead of the lengths, the call returns the bytes the reference
 * writes to <reads>.pseudo_lengths for these reads (compute_ms_pml.cpp:1007-1010): per read '>' id '\n', then every
 * length followed by one space, then '\n'.  ids: the read ids back to back (id i = ids[id_offsets[i] .. id_offsets[i+1]),
 * exactly as BatchLoader extracts them, batch_loader.cpp:104-106).  text: caller buffer of text_capacity bytes;
 * *text_bytes receives the bytes written.  SPB_ERR_LIMIT if the buffer is too small (*text_bytes = bytes needed so
 * far).  stats/opts as in spb_pml_query_host (opts->length_bytes is ignored). */
int spb_pml_query_host_text(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                            const uint8_t* ids, const uint64_t* id_offsets, char* text, uint64_t text_capacity,
                            uint64_t* text_bytes, spb_bin_stats* stats, const spb_query_opts* opts);

/* Same on the gpu_slot-th GPU of a replicated index (0 .. spb_index_device_count - 1).  Calls for DIFFERENT slots of one
 * handle may run concurrently from different host threads (this is how `spumoni run` feeds several GPUs: batches of reads
 * go round robin to one query thread per GPU and are written in input order); calls for the same slot may not. */
int spb_pml_query_host_text_on(spb_index* idx, int gpu_slot, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                               const uint8_t* ids, const uint64_t* id_offsets, char* text, uint64_t text_capacity,
                               uint64_t* text_bytes, spb_bin_stats* stats, const spb_query_opts* opts);

/* The general form: PML or MS (the mode of the index), with or without document numbers -- every output the reference
 * writes as text comes back formatted: <reads>.pseudo_lengths | .lengths, .pointers (MS; the reference's 64-bit values),
 * .doc_numbers (-d, index opened with SPB_OPEN_DOC) (compute_ms_pml.cpp:1003-1010, :1184-1194).  pointers_text must be
 * non-NULL exactly for an MS index; docs_text may be NULL.  On SPB_ERR_LIMIT every buffer's `bytes` holds the bytes needed
 * up to the sub-batch that did not fit (grow and repeat). */
typedef struct {
    char* text;        /* caller buffer */
    uint64_t capacity; /* its size */
    uint64_t bytes;    /* out: bytes written */
} spb_text_buf;
int spb_query_host_text_on(spb_index* idx, int gpu_slot, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                           const uint8_t* ids, const uint64_t* id_offsets, spb_text_buf* lengths_text, spb_text_buf* pointers_text,
                           spb_text_buf* docs_text, spb_bin_stats* stats, const spb_query_opts* opts);

/* Document-number variants: pml_t::matching_statistics(read, len, lengths, doc_nums) compute_ms_pml.cpp:734-737 ->
 * pml_pointers::_query :288-340, and ms_t::matching_statistics(read, len, lengths, pointers, doc_nums) :814-828 ->
 * ms_pointers::_query :625-682.  doc_nums[offsets[i] + j] is the document of position j of read i.  The index must
 * have been opened with SPB_OPEN_DOC; lengths are uint32_t (opts->length_bytes 0 or 4); thread-per-read kernels only. */
int spb_pml_query_host_doc(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                           uint32_t* lengths, uint32_t* doc_nums, spb_bin_stats* stats, const spb_query_opts* opts);
int spb_ms_query_host_doc(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                          uint32_t* lengths, uint64_t* pointers, uint32_t* doc_nums, spb_bin_stats* stats,
                          const spb_query_opts* opts);
int spb_pml_query_device_doc(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                             uint32_t* d_lengths, uint32_t* d_doc_nums, spb_bin_stats* d_stats,
                             const spb_query_opts* opts, void* cuda_stream);
int spb_ms_query_device_doc(spb_index* idx, const uint8_t* d_reads, const uint64_t* d_offsets, uint64_t n_reads,
                            uint32_t* d_lengths, uint64_t* d_pointers, uint32_t* d_doc_nums, spb_bin_stats* d_stats,
                            const spb_query_opts* opts, void* cuda_stream);
/* number of documents in the loaded document array (0 if the index was opened without it) */
uint32_t spb_index_num_docs(const spb_index* idx);

/* ---- classification helpers ---------------------------------------------------------- */
/* EmpNullDatabase::load emp_null_database.cpp:104-110 */
int spb_nulldb_load(const char* path, spb_nulldb* out);
/* Null-database generation (the last stage of `spumoni build`, run through the GPU path):
 * generate_null_pml_statistics / generate_null_ms_statistics compute_ms_pml.cpp:1410-1486 -- every null read is
 * upper-cased and reversed, queried, and its lengths appended to null_stats (offsets[n_reads] values, caller-
 * allocated) -- followed by the EmpNullDatabase constructor emp_null_database.cpp:42-79 (num_values, mean_null_stat,
 * percentile_value = largest value occurring >= 5 times).  ks_stat_threshold is set to 0: the reference derives it
 * from rand()-sampled KS tests (compute_ms_pml.cpp:1549-1663) and never reads it in `run`. */
int spb_nulldb_generate_host(spb_index* idx, const uint8_t* reads, const uint64_t* offsets, uint64_t n_reads,
                             uint32_t* null_stats, spb_nulldb* summary);
/* EmpNullDatabase::serialize emp_null_database.cpp:82-102 (int_vector width max(ceil(log2(max)), 1), values
 * truncated to that width like sdsl does) */
int spb_nulldb_save(const char* path, const spb_nulldb* summary, const uint32_t* null_stats);
/* max_value_thr derivation compute_ms_pml.cpp:871-875 (PML) / :1061-1063 (MS) */
uint64_t spb_max_value_threshold(double percentile_value, int is_pml, int use_dna_letters, int use_promotions);

/* ---- index construction on the GPU ----------------------------------------------------------------------- */
/* What `spumoni build` obtains from its PFP toolchain -- Big-BWT and pfp_thresholds, run as external programs at
 * /root/reference/src/spumoni.cpp:442-545 -- computed on the device from the text itself: per run of the BWT its head, its
 * length, its MONI threshold (the leftmost minimum of LCP between the previous run of the character and this one; 0 for
 * the first) and the SA values at its first and last position.  These are the contents of <ref>.bwt.heads, .bwt.len,
 * .thr_pos, .ssa and .esa, which the reference's own constructors read (compute_ms_pml.cpp:44-82, :357-439,
 * thresholds_ds.hpp:393-426) before serialising the index.
 * text: n bytes = the indexed text followed by its terminator, a zero byte; every other byte must be > 0x01.
 * n < 2^31 - 16 in this version (SPB_ERR_LIMIT otherwise).  The arrays are allocated by the library: spb_build_parts_free. */
typedef struct {
    uint64_t n, r;   /* BWT length, number of runs */
    uint8_t* heads;  /* r */
    uint64_t* lens;  /* r */
    uint64_t* thr;   /* r */
    uint64_t* ssa;   /* r */
    uint64_t* esa;   /* r */
} spb_build_parts;
int spb_build_parts_device(const uint8_t* text, uint64_t n, int device, spb_build_parts* out);
void spb_build_parts_free(spb_build_parts* p);
/* writes <ref_file>.bwt.heads / .bwt.len / .thr_pos / .ssa / .esa with the 5-byte integers the reference reads */
int spb_build_parts_write(const char* ref_file, const spb_build_parts* p);

/* ---- compact transports (host side, no GPU needed) ----------------------------------------------- */
/* bytes of the bit map for n_bases bases (whole 32-byte blocks) / of the packed array (incl. the padding the device reads) */
uint64_t spb_bits_bytes(uint64_t n_bases);
uint64_t spb_packed_bytes(uint64_t n_bases);
/* reads (n_bases bytes, all reads back to back) -> 2 bits per base, base i in bits 2 (i % 4) .. of byte i / 4, A C G T = 0 1 2 3
 * (with SPB_QUERY_TOUPPER in `flags` also a c g t, as the query will upper-case them anyway).  Every other byte is recorded
 * verbatim in the exception list (positions ascending): arbitrary bytes stay legal input.  packed: spb_packed_bytes(n_bases)
 * bytes.  SPB_ERR_LIMIT if exc_capacity is too small (*n_exc = entries needed).  Runs on all host cores (OpenMP). */
int spb_pack_reads(const uint8_t* reads, uint64_t n_bases, uint32_t flags, uint8_t* packed, uint64_t* exc_pos,
                   uint8_t* exc_val, uint64_t exc_capacity, uint64_t* n_exc);
/* bit map (SPB_LENGTH_BITS) -> lengths of `length_bytes` (0/4, 2, 1; clamped like the device would) */
int spb_pml_expand_bits(const void* bits, const uint64_t* offsets, uint64_t n_reads, uint32_t length_bytes, void* lengths);
/* clamped 1- / 2-byte PML lengths -> exact uint32_t lengths */
int spb_pml_widen_lengths(const void* narrow, uint32_t narrow_bytes, const uint64_t* offsets, uint64_t n_reads, uint32_t* lengths);
/* contiguous read ranges balanced by bases: cuts[s] .. cuts[s + 1] is the range of shard s (cuts has n_shards + 1 entries) */
int spb_shard_by_bases(const uint64_t* offsets, uint64_t n_reads, uint32_t n_shards, uint64_t* cuts);
/* test hooks of the host-side planning: sub-batches of a host call (first read of each, then n_reads) and the
 * longest-first processing order of a sub-batch (*sorted = 0: lengths uniform, order not used) */
int spb_plan_chunks(const uint64_t* offsets, uint64_t n_reads, uint64_t want_reads, uint64_t min_bases, uint64_t max_bases,
                    uint64_t* chunk_first_read, uint64_t capacity, uint64_t* n_chunks);
int spb_plan_order(const uint64_t* offsets, uint64_t n_reads, uint32_t* order, int* sorted);

/* ---- utilities ----------------------------------------------------------------------- */
int spb_host_alloc(void** p, size_t bytes); /* pinned host memory (portable: page-locked for every GPU) */
void spb_host_free(void* p);
int spb_device_count(void);
/* number of kernels this library has launched in the calling process (bench `gpu_launches`) */
uint64_t spb_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* SPUMONI_B200_H */
