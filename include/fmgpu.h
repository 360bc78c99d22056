/* fmgpu.h — C ABI of the B200-native batched query engine for index4j FM-indexes.
 *
 * The reference (dynatrace-oss/index4j) has no FFI seam: the drop-in boundary is the public method
 * set of com.dynatrace.fm.FmIndex plus its serialized form.  Every entry point
 * below names the reference method it replaces; paths are relative to
 *   indices/src/main/java/com/dynatrace/   (FM = fm/FmIndex.java, SER = serialization/Serialization.java)
 * A Java host binds these with Panama FFM (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every buffer it passes; the library owns device memory.
 *   - return 0 = OK; < 0 = call-level failure (bad argument, malformed stream, CUDA error); the
 *     message is in fmgpu_last_error() (thread-local).
 *   - per-query Java exceptions are reported in status arrays with the FMGPU_ST_* codes, which map
 *     1:1 onto the reference's exception messages so a Java shim can re-throw the identical exception.
 *   - "chars" are UTF-16 code units exactly as in a Java char[]; pattern i is
 *     chars[pat_off[i] .. pat_off[i+1]).
 *   - *_device variants take DEVICE pointers and a cudaStream_t (as void*), enqueue asynchronously and
 *     do not wait for their work (fmgpu_locate_batch_device synchronizes once to learn the hit total); calls that end up
 *     sharing internal scratch buffers are ordered on the device by events, whatever streams they were given.  The
 *     host-pointer variants copy H2D, run, copy D2H and synchronize.
 *   - there is no CPU fallback: without a CUDA device every call fails with FMGPU_ERR_CUDA.
 */
#ifndef FMGPU_H
#define FMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMGPU_OK 0
#define FMGPU_ERR_ARG (-1)     /* null handle / bad argument */
#define FMGPU_ERR_FORMAT (-2)  /* malformed stream, incl. "Incompatible serial versions! ..." (SER:46-56) */
#define FMGPU_ERR_CUDA (-3)    /* CUDA runtime failure (no device, out of memory, launch error) */
#define FMGPU_ERR_CAPACITY (-4) /* caller's output buffer too small (locate positions_cap) */
#define FMGPU_ERR_UNSUPPORTED (-5)

/* Per-query status = the Java exception the reference would have thrown. */
#define FMGPU_ST_OK 0
#define FMGPU_ST_NOT_ENABLED 1    /* RuntimeException("Text recovery not enabled at build time")   FM:566,611 */
#define FMGPU_ST_POS_NEGATIVE 2   /* RuntimeException("Requested position less than 0")            FM:570,615 */
#define FMGPU_ST_STOP_TOO_LONG 3  /* RuntimeException("Stop position longer than index string")    FM:574 */
#define FMGPU_ST_POS_TOO_LONG 4   /* RuntimeException("Requested position longer than index string") FM:619 */
#define FMGPU_ST_DST_TOO_SMALL 5  /* RuntimeException("Supplied destination is not large enough")  FM:591 */
#define FMGPU_ST_DST_ZERO 6       /* IllegalArgumentException("Supplied destination for extraction has size zero") FM:623 */
#define FMGPU_ST_NO_BOUNDARY 7    /* IllegalArgumentException("Boundary does not exist")           FM:659,792,849 */
#define FMGPU_ST_DOES_NOT_FIT 8   /* RuntimeException("Extraction does not fit in the supplied destination. Currently extracted: N") FM:733,817,894; N in len_out */
#define FMGPU_ST_RRR_RANGE 11     /* IllegalArgumentException("Out of range access. Requested P when range is [0, L)") RRR:316-323 */
#define FMGPU_ST_CHAR_EXCEEDS 10  /* RuntimeException("Found a character that exceeds (32767): it was N") FM:261-267 (UTF-8 entry points; N in counts_out) */
#define FMGPU_ST_NO_TERMINATION 12 /* the reference never returns: an LF walk of locate (FM:531-537) has run for more than
                                    * getInputLength() steps, i.e. in a cycle — only possible on an index with more than 256
                                    * symbols where inverseSelect truncated a single-symbol block's symbol (WF:1329-1332) */
#define FMGPU_ST_INDEX_OOB 9      /* ArrayIndexOutOfBoundsException (empty pattern FM:456; rank(size,.) on a superblock boundary, wavelet/WaveletFixedBlockBoosting.java:1022-1026) */

#define FMGPU_MODE_BOTH 0  /* FmIndex.extractUntilBoundary       FM:640 */
#define FMGPU_MODE_LEFT 1  /* FmIndex.extractUntilBoundaryLeft   FM:772 */
#define FMGPU_MODE_RIGHT 2 /* FmIndex.extractUntilBoundaryRight  FM:844 */

typedef struct fmgpu_index fmgpu_index;

typedef struct fmgpu_opts {
    int32_t device;         /* CUDA device ordinal (used when n_devices == 0); -1 = current device */
    int32_t host_threads;   /* threads used to re-lay the index out at load; 0 = all cores */
    int32_t n_devices;      /* > 0: replicate the index on devices[0 .. n_devices); < 0: on every visible device; 0: `device` alone */
    int32_t locate_sample_rate; /* device-side denser sampling of the SA rows for locate (fmgpu_set_locate_dense): > 0 requested rate,
                                 * 0 = default (FMGPU_LOCATE_SAMPLE_RATE in the environment, else 8), < 0 = none */
    const int32_t* devices; /* device ordinals, distinct */
    uint64_t reserved[1];
} fmgpu_opts;

const char* fmgpu_last_error(void);
const char* fmgpu_version(void);

/* Replaces Serialization.readFromByteArray(FmIndex::read, bytes) (SER:89-100, FM:983-1025).
 * Accepts the ObjectOutputStream framing (AC ED 00 05 + block-data records) or the bare
 * DataOutput primitives.  Parses the stream, re-lays the structures out into the device format
 *  and uploads them once.  opts may be NULL.
 *
 * The handle is the reference's one immutable, @ThreadSafe FmIndex (FM:82): with a device list in opts the layout is uploaded
 * to the first device and copied to the others over NVLink (cudaMemcpyPeer); every host-pointer batch call then cuts the
 * caller's batch into one contiguous slice per device, runs the slices at once and writes disjoint ranges of the caller's
 * outputs.  Any number of host threads may call into one handle concurrently (each call leases its own streams and scratch
 * buffers on the devices it uses).  The *_device entry points run on the replica whose device holds the caller's buffers. */
int fmgpu_index_load_serialized(const uint8_t* buf, size_t len, const fmgpu_opts* opts, fmgpu_index** out);
void fmgpu_index_free(fmgpu_index* idx);

int32_t fmgpu_input_length(const fmgpu_index* idx);     /* FmIndex.getInputLength()    FM:929 (= n+1) */
int32_t fmgpu_alphabet_length(const fmgpu_index* idx);  /* FmIndex.getAlphabetLength() FM:939 */
int32_t fmgpu_sample_rate(const fmgpu_index* idx);
int32_t fmgpu_extract_enabled(const fmgpu_index* idx);
int32_t fmgpu_device(const fmgpu_index* idx);             /* the first (primary) device */
int32_t fmgpu_num_devices(const fmgpu_index* idx);        /* replicas */
int32_t fmgpu_device_at(const fmgpu_index* idx, int32_t i);
uint64_t fmgpu_device_bytes(const fmgpu_index* idx);    /* bytes of HBM held by the index */
/* component sizes (bytes): [0] cells [1] level records [2] node records [3] block descriptors
 * [4] occurrence records [5] sampled-row groups+offsets [6] SA samples [7] ISA samples */
void fmgpu_layout_bytes(const fmgpu_index* idx, uint64_t out8[8]);

/* Page-locked host memory.  The host-pointer batch calls copy with cudaMemcpyAsync, which only overlaps with the kernels (and
 * reaches the PCIe rate) from page-locked buffers: register the caller's arrays once (a Java host: the MemorySegments of its
 * Arena) or allocate them here.  Pageable buffers still work, at a fraction of the rate. */
int fmgpu_host_register(void* p, size_t bytes);
int fmgpu_host_unregister(void* p);
int fmgpu_host_alloc(size_t bytes, void** out);
int fmgpu_host_free(void* p);

/* FmIndex.count(char[] p, int off, int len)  FM:455-474 — one result per pattern.
 * status_out may be NULL.
 * Transport of the host-pointer call: it is bound by the upload of the chars (2 bytes each), so on a single-device handle a pool
 * of host threads narrows every chunk of the batch whose chars all fit a byte (Latin-1: log text) to bytes + chunk-relative
 * uint32 offsets into the library's own page-locked staging buffer while earlier chunks are on the wire, and the device widens
 * them again; other chunks go as they are.  The caller's arrays need not be page-locked for the packed chunks.  Pool size =
 * fmgpu_host_pack_threads(): half the hardware threads (divided by LOCAL_WORLD_SIZE), at most 8, none below 8 — then, and with
 * FMGPU_HOST_PACK=0, every chunk goes as it is; FMGPU_PACK_THREADS overrides the size. */
int32_t fmgpu_host_pack_threads(void);
int fmgpu_count_batch(fmgpu_index* idx, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat,
                      int32_t* counts_out, int32_t* status_out);
int fmgpu_count_batch_device(fmgpu_index* idx, const uint16_t* d_chars, const uint64_t* d_pat_off, uint64_t total_chars,
                             uint32_t n_pat, int32_t* d_counts_out, int32_t* d_status_out, void* cuda_stream);

/* q-gram start table of the backward search: at load the search kernel computes, for every q-gram of alphabet codes (q = the
 * largest value with sigma^q <= min(2^25, 4 x text length) entries of 8 bytes, q >= 2; none for alphabets above 5792 symbols), the SA range after its q chars —
 * the state of FmIndex.count (FM:455-474) after q - 1 steps of its loop — and count / locate start every pattern of >= q
 * known chars from that entry instead of from C[] of its last char.  Results are identical (the entries ARE the loop's states;
 * q-grams on whose way the reference throws are not entered).  fmgpu_set_start_table(idx, 0) makes every pattern start from its
 * last char (the work counters of fmgpu_last_stats then count every step); FMGPU_START_TABLE=0 in the environment skips
 * building it, FMGPU_START_TABLE_LOG2=n caps it at 2^n entries.  fmgpu_start_table_q: q, 0 = no table. */
int fmgpu_set_start_table(fmgpu_index* idx, int enable);
int32_t fmgpu_start_table_q(const fmgpu_index* idx);

/* Denser SA sampling on the device.  FmIndex keeps suffixes[] for the rows whose suffix starts at a multiple of sampleRate
 * (FM:343-357) and locate() LF-walks every hit to the nearest such row (FM:531-537): (sampleRate - 1) / 2 steps per hit, the
 * size / speed trade of the serialized index.  HBM makes a different trade affordable: at load the LF kernels walk the text once
 * (2 x length LF steps, ~0.1 s per 2^30 chars) and also record the rows of every multiple of a smaller rate d (the largest
 * divisor of sampleRate <= fmgpu_opts.locate_sample_rate), as a plain mark vector with rank counters + their positions
 * (length / 7 + 4 * length / d bytes).  Walks then end after (d - 1) / 2 steps.  Positions, their order and statuses are
 * unchanged (position = sampled position + distance walked, whichever sample ends the walk); on indexes where an LF step of
 * the reference can throw or cycle (length % 2^20 == 0; more than 256 symbols) no dense samples are built, so the reference's
 * full walk and its exception are what run.  fmgpu_set_locate_dense(idx, 0) makes locate use the index's own samples
 * (A/B measurements, tests); fmgpu_locate_sample_rate = the rate in effect (sampleRate when none were built). */
int fmgpu_set_locate_dense(fmgpu_index* idx, int enable);
int32_t fmgpu_locate_sample_rate(const fmgpu_index* idx);
uint64_t fmgpu_dense_sample_bytes(const fmgpu_index* idx); /* HBM held by the dense marks + positions (part of fmgpu_device_bytes) */

/* UTF-8 byte patterns: FmIndex.convertBytePatternToCharPattern(byte[] p, 0, p.length, dst) FM:239-298 followed by
 * count(dst, 0, n) / locate(dst, 0, n, ...).  Pattern i is the byte[] bytes[pat_off[i], pat_off[i+1]) — pat_off are BYTE
 * offsets.  The bytes (1 per char for ASCII / Latin-1 logs instead of the 2 of a char[]) cross PCIe and are decoded on
 * the device with the reference's branch structure.  Status FMGPU_ST_CHAR_EXCEEDS (offending code point in counts_out)
 * where the converter throws; FMGPU_ST_INDEX_OOB where a multi-byte sequence runs past the end of its byte[]
 * (ArrayIndexOutOfBoundsException) and, as for char[] patterns, for an empty pattern. */
int fmgpu_count_batch_utf8(fmgpu_index* idx, const uint8_t* bytes, const uint64_t* pat_off, uint32_t n_pat,
                           int32_t* counts_out, int32_t* status_out);
int fmgpu_count_batch_utf8_device(fmgpu_index* idx, const uint8_t* d_bytes, const uint64_t* d_pat_off, uint64_t total_bytes,
                                  uint32_t n_pat, int32_t* d_counts_out, int32_t* d_status_out, void* cuda_stream);
int fmgpu_locate_batch_utf8(fmgpu_index* idx, const uint8_t* bytes, const uint64_t* pat_off, uint32_t n_pat, int32_t max_hits,
                            int32_t* n_hits_out, uint64_t* hit_off_out, int32_t* positions_out, uint64_t positions_cap,
                            int32_t* status_out);

/* FmIndex.locate(char[] p, int off, int len, int[] out, int max)  FM:504-552.
 * max_hits <= 0 means unlimited (FM:544).  Hits of pattern i are written to
 * positions_out[hit_off_out[i] .. hit_off_out[i+1]) in SA-row order (the order Java fills its
 * array).  Two-phase sizing: with positions_out == NULL only n_hits_out / hit_off_out are produced.
 * If positions_cap < total hits the call returns FMGPU_ERR_CAPACITY (hit_off_out is still valid).
 * Where an LF walk of a hit indexes outside the reference's arrays (rank(size, .) on a superblock boundary, WF:1022-1026)
 * the reference throws out of locate(): status_out[pattern] = FMGPU_ST_INDEX_OOB (FMGPU_ST_NO_TERMINATION for a walk in a
 * cycle) and that hit's slot holds -1; n_hits_out keeps the number of slots. */
int fmgpu_locate_batch(fmgpu_index* idx, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, int32_t max_hits,
                       int32_t* n_hits_out, uint64_t* hit_off_out, int32_t* positions_out, uint64_t positions_cap,
                       int32_t* status_out);
/* Device form: d_positions_out has room for positions_cap entries; *total_hits_out (host) receives
 * the number of hits (this call synchronizes the stream once to learn it). */
int fmgpu_locate_batch_device(fmgpu_index* idx, const uint16_t* d_chars, const uint64_t* d_pat_off, uint64_t total_chars,
                              uint32_t n_pat, int32_t max_hits, int32_t* d_n_hits_out, uint64_t* d_hit_off_out,
                              int32_t* d_positions_out, uint64_t positions_cap, int32_t* d_status_out,
                              uint64_t* total_hits_out, void* cuda_stream);

/* FmIndex.extract(int start, int stop, char[] destination, int offset)  FM:564-608 with destination = the slot
 * arena[arena_off[i] .. arena_off[i+1]) (its length is destination.length) and the same `offset` (>= 0) for every item:
 * the chars land at slot + offset, "Supplied destination is not large enough" when destination.length - offset < stop - start
 * (FM:591).  len_out[i] = stop - start on success. */
int fmgpu_extract_batch(fmgpu_index* idx, const int32_t* start, const int32_t* stop, uint32_t n, uint16_t* arena,
                        const uint64_t* arena_off, int32_t offset, int32_t* len_out, int32_t* status_out);
int fmgpu_extract_batch_device(fmgpu_index* idx, const int32_t* d_start, const int32_t* d_stop, uint32_t n, uint16_t* d_arena,
                               const uint64_t* d_arena_off, int32_t offset, int32_t* d_len_out, int32_t* d_status_out, void* cuda_stream);

/* FmIndex.extractUntilBoundary / ...Left / ...Right (int from, char[] destination, int offset, char boundary) (FM:640-922) with
 * destination = new char[dst_len] = the slot arena[i*dst_len .. (i+1)*dst_len) and the same `offset` (>= 0) for every item.  As in
 * the reference the record lands at slot + offset, the left walk may use the WHOLE destination as scratch (remaining =
 * destination.length, FM:662) and `offset` enters the "does not fit" test and its N (FM:732-737, :817-821, :894-898) but not
 * the returned length; a left part that does not fit behind `offset` makes System.arraycopy throw (FM:688-690): status
 * FMGPU_ST_INDEX_OOB.  len_out[i] = returned length (or N of the "does not fit" message when status is
 * FMGPU_ST_DOES_NOT_FIT).  Only arena[i*dst_len + offset, +len) is defined, like the Java array beyond the returned length.
 * (from == getInputLength() - 1, the terminator's own position: the reference's end-of-text rule, FM:745-752, returns a length
 * one beyond the chars it wrote; that last slot keeps what the destination held, in Java and here.) */
int fmgpu_extract_until_boundary_batch(fmgpu_index* idx, const int32_t* from, uint32_t n, uint16_t boundary, int32_t dst_len,
                                       int32_t offset, int32_t mode, uint16_t* arena, int32_t* len_out, int32_t* status_out);
int fmgpu_extract_until_boundary_batch_device(fmgpu_index* idx, const int32_t* d_from, uint32_t n, uint16_t boundary,
                                              int32_t dst_len, int32_t offset, int32_t mode, uint16_t* d_arena, int32_t* d_len_out,
                                              int32_t* d_status_out, void* cuda_stream);

/* Fused locate -> extractUntilBoundary — the reference's "extracting whole records" flow (README.md:98-107,
 * jmh/.../FmIndexThroughputBenchmark.java:231-249):
 *     found = fmi.locate(pattern, 0, len, locations, max);
 *     for (i < found) length = fmi.extractUntilBoundary(locations[i], destination = new char[dst_len], 0, boundary);
 * with every DISTINCT record read from the index once: hits that lie in the same record (of one pattern or of different
 * patterns of the batch) share it.  Per hit h: rec_index_out[h] = row of its record in rec_arena (row u = rec_arena[u*dst_len ..
 * (u+1)*dst_len)), or -1 when the call has no record for it (status != 0); len_out[h] / status_out[h] = exactly what the
 * reference's call returns or throws for THAT hit (the "does not fit" test works on 4-char chunks counted from the hit, so
 * it can differ between two hits of one record, FM:732-737); the hit's record is rec_arena[row][0 .. len_out[h]).
 * rec_cap = rows the arena has room for (n hits always suffice); *n_records_out = distinct records.  The records keep no
 * particular order.
 *   fmgpu_extract_records_batch[_device]: the hits are given (text positions, e.g. of an earlier locate);
 *   fmgpu_locate_records_batch: patterns in, hits + records out; with positions_out == NULL only n_hits_out / hit_off_out /
 *   pat_status_out are produced (sizing pass: hits_cap must be >= hit_off_out[n_pat]).  These calls run on the primary device. */
int fmgpu_extract_records_batch(fmgpu_index* idx, const int32_t* from, uint32_t n, uint16_t boundary, int32_t dst_len,
                                int32_t* rec_index_out, int32_t* len_out, int32_t* status_out, uint16_t* rec_arena, uint64_t rec_cap,
                                uint64_t* n_records_out);
int fmgpu_extract_records_batch_device(fmgpu_index* idx, const int32_t* d_from, uint32_t n, uint16_t boundary, int32_t dst_len,
                                       int32_t* d_rec_index_out, int32_t* d_len_out, int32_t* d_status_out, uint16_t* d_rec_arena,
                                       uint64_t rec_cap, uint64_t* n_records_out, void* cuda_stream);
int fmgpu_locate_records_batch(fmgpu_index* idx, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, int32_t max_hits,
                               uint16_t boundary, int32_t dst_len, int32_t* n_hits_out, uint64_t* hit_off_out, int32_t* pat_status_out,
                               int32_t* positions_out, int32_t* rec_index_out, int32_t* len_out, int32_t* status_out, uint64_t hits_cap,
                               uint16_t* rec_arena, uint64_t rec_cap, uint64_t* n_records_out);

/* The index's wavelet structure (FmIndex.waveletFixedBlockBoosting) queried directly — the reference's public
 * WaveletFixedBlockBoosting.rank(long position, short symbol) WF:1010-1285 and inverseSelect(long position) WF:1305-1537
 * over alphabet CODES (FmIndex maps chars to codes by first appearance, FM:396-435).  out[i] of inverse_select is the
 * packed long Java returns: (rank << 32) | symbol, the bare symbol for position 0.  Status FMGPU_ST_INDEX_OOB where the
 * reference indexes out of its arrays (negative arguments, rank(size, .) on a superblock boundary, inverseSelect outside
 * [0, size)). */
/* fmgpu_wavelet_load_serialized: a handle from a bare WaveletFixedBlockBoosting stream (WaveletFixedBlockBoosting.write WF:1544 /
 * read WF:286) — valid for the two fmgpu_wavelet_* calls and the getters only (symbols are the structure's own short values).
 * fmgpu_rrr_load_serialized: a handle from a bare RrrVector stream (RrrVector.write / read, bitsequence/RrrVector.java:430-469)
 * — valid for fmgpu_rrr_rank_access_batch only.  Every other call on such handles returns FMGPU_ERR_UNSUPPORTED.  Free both
 * with fmgpu_index_free. */
int fmgpu_wavelet_load_serialized(const uint8_t* buf, size_t len, const fmgpu_opts* opts, fmgpu_index** out);
int fmgpu_rrr_load_serialized(const uint8_t* buf, size_t len, const fmgpu_opts* opts, fmgpu_index** out);
/* RrrVector.rankOnes(int) RRR:358-396 and access(int) RRR:314-349 per position (on an RRR handle, or the sampledSuffixes
 * vector of an FmIndex handle).  rankOnes: position < 0 -> 0, position >= length -> total ones; access outside [0, length)
 * throws in the reference: status_out FMGPU_ST_RRR_RANGE, access_out 0.  rankZeroes(p) = p - rankOnes(p) (RRR:405). */
int fmgpu_rrr_rank_access_batch(fmgpu_index* idx, const int32_t* pos, uint32_t n, int32_t* rank_ones_out, int32_t* access_out,
                                int32_t* status_out);
int fmgpu_wavelet_rank_batch(fmgpu_index* idx, const int64_t* pos, const int32_t* sym, uint32_t n, int64_t* out, int32_t* status_out);
int fmgpu_wavelet_inverse_select_batch(fmgpu_index* idx, const int64_t* pos, uint32_t n, int64_t* out, int32_t* status_out);

/* Index PRODUCTION, device stage (not a query-path call): from the suffix array of (alphabet codes + sentinel),
 * both already in device memory, to the pieces the FmIndex constructor derives from it (FM:343-394): the BWT
 * (d_bwt_out[length]), the sampled-row marks (one bit per row, LSB first, d_mask_words_out[(length+31)/32]), the SA samples in
 * row order (d_suffixes_out, *n_sampled_out of them) and the inverse-SA samples with the wrap entry (d_positions_out[length /
 * sample_rate + 2], may be NULL).  Synchronizes the stream.  The host-side producer (index4j_b200/csrc/host) encodes and
 * serializes them. */
int fmgpu_build_bwt_samples_device(const uint16_t* d_codes, const int32_t* d_sa, int64_t length, int32_t sample_rate, uint16_t* d_bwt_out,
                                   uint32_t* d_mask_words_out, int32_t* d_suffixes_out, int64_t suffixes_cap, int32_t* d_positions_out,
                                   int64_t* n_sampled_out, void* cuda_stream);

/* Sharded index — texts beyond the reference's 2^31-char limit (int length / char[] input, FM:131,155,335-341): one FmIndex per
 * GPU over a text shard that also holds the first max_pattern_len - 1 chars of the next shard.  The reference has no such
 * mode; each shard's results are bit-exact with the Java FmIndex of that shard, and these kernels compose them: a hit is OWNED
 * by the shard in which it starts before owned_len, at most max_hits hits per pattern survive globally — lowest shard first,
 * SA order inside a shard (max_hits <= 0: all).  Device pointers, the caller's stream, no synchronization; the host
 * (index4j_b200/sharded.py, one process per GPU) runs the two NCCL exchanges between them.
 *   keep : d_kept[p] = owned hits of pattern p among the local hits d_pos[d_hit_off[p] .. d_hit_off[p+1]), capped at max_hits
 *          -> all-gather of d_kept over the ranks = d_all_kept[world][n_pat]
 *   plan : d_take[r][p] = what rank r contributes after the global cut, d_n_hits[p] / d_hit_off[n_pat + 1] = global hits per
 *          pattern and their offsets, d_roff[r][n_pat + 1] = offsets inside rank r's contribution, d_totals[world + 1] = size of
 *          every rank's contribution + the merged total, d_rank_base[world] = where rank r's contribution starts in the receive buffer
 *   pack : this rank's contribution as global positions (local + text_start, int64) in pattern order, d_send[d_totals[rank]]
 *          -> exchange: every rank receives every contribution, rank r's at d_recv + d_rank_base[r]
 *   merge: d_out[d_hit_off[p] ..] = the contributions of pattern p in rank order */
int fmgpu_shard_keep_device(const int32_t* d_pos, const uint64_t* d_hit_off, uint32_t n_pat, int32_t owned_len, int32_t max_hits,
                            int32_t* d_kept, void* cuda_stream);
int fmgpu_shard_plan_device(const int32_t* d_all_kept, uint32_t n_pat, uint32_t world, uint32_t rank, int32_t max_hits, int32_t* d_take,
                            int32_t* d_n_hits, uint64_t* d_hit_off, uint64_t* d_roff, uint64_t* d_totals, uint64_t* d_rank_base,
                            void* cuda_stream);
int fmgpu_shard_pack_device(const int32_t* d_pos, const uint64_t* d_hit_off, uint32_t n_pat, int32_t owned_len, int64_t text_start,
                            const int32_t* d_take_mine, const uint64_t* d_send_off, int64_t* d_send, void* cuda_stream);
int fmgpu_shard_merge_device(const int64_t* d_recv, const uint64_t* d_rank_base, const uint64_t* d_roff, const int32_t* d_take,
                             const uint64_t* d_out_off, uint32_t n_pat, uint32_t world, int64_t* d_out, void* cuda_stream);

/* Work counters of the most recent batch call on this index (device-side counted, read back here):
 * [0] rank queries that touched memory  [1] wavelet levels walked by rank queries
 * [2] LF steps (inverseSelect walks)     [3] wavelet levels walked by LF steps
 * [4] sampled-row bit tests              [5] kernels launched by the call
 * [6] 32-byte records actually loaded by the backward-search kernel (cells + occurrence records; two positions of one
 *     pattern that fall into the same cell / record share one load)
 * [7] records beyond the cell / descriptor the queries need: occurrence records of rank queries (0 or 1 each), level records
 *     of LF steps (one per TWO wavelet levels)
 * Used by bench.py for the roofline's algorithmic-bytes figure. */
int fmgpu_last_stats(fmgpu_index* idx, uint64_t out8[8]);
/* The kernels keep these counters only while enabled (default: off — the production instantiations carry none;
 * [5], the launch count, is always kept). */
int fmgpu_set_stats(fmgpu_index* idx, int enable);
/* Same, all FMGPU_N_STATS counters (n_out >= FMGPU_N_STATS):
 * [8 + k] rank queries of the backward-search kernel whose (block, symbol) cell had kind k: 1 = CONST (symbol absent: the
 *     cell is the answer), 2 = RUN (single-symbol block), 3 = THROW, 5 = position list (<= 14 occurrences), 4 / 7 = position
 *     lists over 1024- / 4096-position ranges, 6 = bit vector; kinds 4-7 fetch one record. */
#define FMGPU_N_STATS 16
int fmgpu_last_stats_ex(fmgpu_index* idx, uint64_t* out, uint32_t n_out);

/* Measurement hooks (bench.py): with timing enabled every count/locate call brackets its backward-search
 * kernel with CUDA events on the stream it is launched on; fmgpu_search_kernel_ms returns the device
 * time of the kernel launched `calls_back` calls ago (0 = the most recent; up to 64 are kept). */
int fmgpu_set_timing(fmgpu_index* idx, int enable);
int fmgpu_search_kernel_ms(fmgpu_index* idx, uint32_t calls_back, float* ms_out);
/* The same for each of the three dominant kernels, on replica `device_index` (0 = the primary device): the launch of that
 * kind `calls_back` launches ago. */
#define FMGPU_KERNEL_COUNT 0   /* k_count:   backward search */
#define FMGPU_KERNEL_LOCATE 1  /* k_locate:  LF walks of locate */
#define FMGPU_KERNEL_EXTRACT 2 /* k_extract: LF walks of extract / extractUntilBoundary* */
int fmgpu_kernel_ms(fmgpu_index* idx, int32_t kind, uint32_t device_index, uint32_t calls_back, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* FMGPU_H */
