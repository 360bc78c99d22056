// libfmhost: host-side index PRODUCER and synthetic-workload generator (C ABI).
//
// Role in this repo: the GPU query engine (libfmgpu) consumes a *serialized* FmIndex in the
// reference's own Serialization layout.  The reference builds that on the JVM
// (indices/src/main/java/com/dynatrace/fm/FmIndex.java:155-174, write :948-975); no JVM exists in
// this image, so this library produces equivalent streams natively for tests and benchmarks.
// It is NOT on the query path and NOT the oracle; the oracle (oracle/) and the GPU library parse
// the stream independently.
//
// What "equivalent" means: every field of the stream grammar is produced by the
// same rules as the reference constructor — first-appearance alphabet map (:396-435), cumulative
// counts (:307-327), suffix-array sampling and sampled-row bitvector (:329-372), BWT (:374-394),
// WFBB over the BWT with rrr sample rate = sampleRate (:173).  Byte identity with a JVM-written
// stream is attempted (HashMap key order emulation, ObjectOutputStream block-data framing) but
// cannot be verified here.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "packed.hpp"
#include "rrr_enc.hpp"
#include "sais.hpp"
#include "wfbb_enc.hpp"

using namespace fmhost;

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------------------------
// java.io.ObjectOutputStream-compatible sink: big-endian primitives, optional block-data framing
// (stream header AC ED 00 05; records 0x77 <u8 len> / 0x7A <i32 len>, 1024-byte block buffer).
// ---------------------------------------------------------------------------------------------
struct JavaSink {
    std::vector<uint8_t> out;
    bool framed;
    uint8_t buf[1024];
    int pos = 0;
    explicit JavaSink(bool framed_) : framed(framed_) {
        if (framed) {
            const uint8_t hdr[4] = {0xAC, 0xED, 0x00, 0x05};
            out.insert(out.end(), hdr, hdr + 4);
        }
    }
    void drain() {
        if (pos == 0) return;
        if (pos <= 0xFF) {
            out.push_back(0x77);
            out.push_back((uint8_t)pos);
        } else {
            out.push_back(0x7A);
            out.push_back((uint8_t)(pos >> 24));
            out.push_back((uint8_t)(pos >> 16));
            out.push_back((uint8_t)(pos >> 8));
            out.push_back((uint8_t)pos);
        }
        out.insert(out.end(), buf, buf + pos);
        pos = 0;
    }
    inline void u8(uint8_t b) {
        if (!framed) {
            out.push_back(b);
            return;
        }
        if (pos >= 1024) drain();
        buf[pos++] = b;
    }
    void i16(int32_t v) {
        u8((uint8_t)(v >> 8));
        u8((uint8_t)v);
    }
    void i32(int32_t v) {
        u8((uint8_t)(v >> 24));
        u8((uint8_t)(v >> 16));
        u8((uint8_t)(v >> 8));
        u8((uint8_t)v);
    }
    void i64(int64_t v) {
        for (int s = 56; s >= 0; s -= 8) u8((uint8_t)((uint64_t)v >> s));
    }
    // n big-endian longs; the unframed stream takes them in one resize + byte-swapped stores (the word arrays of the packed
    // vectors are most of a serialized index)
    void i64_array(const uint64_t* w, size_t n) {
        if (framed) {
            for (size_t i = 0; i < n; ++i) i64((int64_t)w[i]);
            return;
        }
        const size_t at = out.size();
        out.resize(at + 8 * n);
        uint8_t* p = out.data() + at;
        for (size_t i = 0; i < n; ++i) {
            const uint64_t be = __builtin_bswap64(w[i]);
            memcpy(p + 8 * i, &be, 8);
        }
    }
    void finish() {
        if (framed) drain();
    }
};

void write_intvec(JavaSink& s, const IntVec& v) {  // IntVector.java:196-203
    s.u8(0);
    s.i32(v.length);
    s.i32(v.width);
    s.i64_array(v.data.data(), v.data.size());
}
void write_rrr(JavaSink& s, const RrrEnc& r) {  // RrrVector.java:430-440
    s.u8(0);
    s.i32(r.sample_size);
    s.i32(r.length);
    s.i32(r.total_ones);
    s.i32(r.bits_per_offset_position);
    write_intvec(s, r.classes);
    s.u8(0);  // VariableWidthIntVector.java:175-181
    s.i32((int32_t)r.offsets.size());
    s.i64_array(r.offsets.data(), r.offsets.size());
    write_intvec(s, r.sampled_offset_pos);
    write_intvec(s, r.prefix_sums);
}
void write_wfbb(JavaSink& s, const WfbbEnc& w) {  // WaveletFixedBlockBoosting.java:1544-1570
    s.u8(0);
    s.i64(w.size);
    s.i32(w.sigma);
    s.i32(w.rrr_rate);
    s.i32((int32_t)w.count.size());
    for (int64_t v : w.count) s.i64(v);
    s.i32((int32_t)w.hyper_rank.size());
    for (int64_t v : w.hyper_rank) s.i64(v);
    s.i32((int32_t)w.sb_rank.size());
    for (int32_t v : w.sb_rank) s.i32(v);
    s.i32((int32_t)w.global_mapping.size());
    for (int16_t v : w.global_mapping) s.i16(v);
    s.i32((int32_t)w.sbs.size());
    for (const SuperBlock& S : w.sbs) {  // :1651-1667
        s.i16(S.sigma_m1);
        s.i16(S.block_size_log);
        write_rrr(s, S.rank);
        s.i32((int32_t)S.blocks.size());
        for (const BlockHeader& H : S.blocks) {  // :1607-1613
            s.i32(H.bv_rank);
            s.i32(H.bv_offset);
            s.i32(H.var_off);
            s.i16(H.sigma_m1);
            s.i16(H.tree_height);
        }
        s.i32((int32_t)S.var.size());
        for (uint8_t b : S.var) s.u8(b);
        s.i32((int32_t)S.mapping.size());
        for (int16_t v : S.mapping) s.i16(v);
    }
}

// Iteration order of a java.util.HashMap<Integer,Short> holding these keys (all < 65536, so
// hash == key): table capacity doubles from 16 while size > 0.75*capacity; buckets are visited
// in index order, entries inside a bucket in insertion order (resize splits preserve it).
std::vector<int32_t> hashmap_key_order(const std::vector<int32_t>& insertion_order) {
    size_t cap = 16;
    while (insertion_order.size() > cap * 3 / 4) cap <<= 1;
    std::vector<int32_t> keys = insertion_order;
    std::stable_sort(keys.begin(), keys.end(), [cap](int32_t a, int32_t b) {
        return ((size_t)a & (cap - 1)) < ((size_t)b & (cap - 1));
    });
    return keys;
}

struct BuildResult {
    std::vector<uint8_t> bytes;
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Pieces of the constructor computed elsewhere (on the GPU, fmgpu_build_bwt_samples_device): the BWT over alphabet codes, the
// sampled-row marks (one bit per row, LSB first), the SA samples in row order and the inverse-SA samples incl. the wrap entry.
struct BuildParts {
    const uint16_t* bwt = nullptr;
    const uint32_t* mask_words = nullptr;
    const int32_t* suffixes = nullptr;
    int64_t n_suffixes = 0;
    const int32_t* positions = nullptr;  // length / sampleRate + 2 entries (enable_extract)
};

// Build from text; if sa_in != nullptr it must be the suffix array of (mapped text + sentinel); with `parts` neither a suffix
// array nor the text passes over it are needed.
int build_index(const uint16_t* text, int64_t n, int sample_rate, int enable_extract, int framed,
                int threads, const int32_t* sa_in, int verbose, std::vector<uint8_t>& out_bytes, const BuildParts* parts = nullptr) {
    if (n < 1) {
        g_err = "text must have at least one char";
        return -1;
    }
    if (n + 1 > 0x7fffffffLL) {
        g_err = "text longer than Java's int range";
        return -1;
    }
    if (sample_rate < 1) {
        g_err = "sampleRate must be >= 1";
        return -1;
    }
    double t0 = now_s();
    const int64_t length = n + 1;
    // --- alphabet map in order of first appearance (FmIndex.java:396-435)
    std::vector<int32_t> code_of(65536, -1);
    bool has_nul = false;
    for (int64_t i = 0; i < n; ++i)
        if (text[i] == 0) {
            has_nul = true;
            break;
        }
    std::vector<int32_t> lookup;  // code -> char
    std::vector<int32_t> insertion_keys;
    int32_t next_code = 0;
    if (has_nul) {  // sentinel keeps code 0; the text's own '\0' gets code 1
        lookup.push_back(0);
        next_code = 1;
    }
    code_of[0] = next_code;
    if ((int32_t)lookup.size() <= next_code) lookup.resize(next_code + 1, 0);
    lookup[next_code] = 0;
    insertion_keys.push_back(0);
    ++next_code;
    for (int64_t i = 0; i < n; ++i) {
        uint16_t c = text[i];
        if (code_of[c] < 0) {
            code_of[c] = next_code++;
            lookup.push_back((int32_t)c);
            insertion_keys.push_back((int32_t)c);
        }
    }
    const int32_t n_codes = next_code;  // codes in use: 0 .. n_codes-1
    // monotonicLookUp has alphabet.size()+1 entries (:411): one unused trailing entry when the
    // text itself holds no '\0' (the HashSet already contains the sentinel's '\0').
    if (!has_nul) lookup.push_back(0);
    if ((int64_t)insertion_keys.size() > 32767) {
        g_err = "Input has more than 32767 different symbols";
        return -1;
    }
    std::vector<uint16_t> mapped;
    if (!parts) {
        mapped.resize((size_t)length);
        for (int64_t i = 0; i < n; ++i) mapped[(size_t)i] = (uint16_t)code_of[text[i]];
        mapped[(size_t)n] = 0;
    }

    // --- cumulative counts (:307-327)
    const int32_t n_lookup = (int32_t)lookup.size();
    std::vector<int32_t> C((size_t)n_lookup + 1, 0);
    {
        std::vector<int64_t> cnt((size_t)n_lookup, 0);
        if (parts) {
            for (int64_t i = 0; i < n; ++i) cnt[(size_t)code_of[text[i]]]++;
            cnt[0]++;  // the sentinel
        } else {
            for (int64_t i = 0; i < length; ++i) cnt[mapped[(size_t)i]]++;
        }
        int64_t acc = 0;
        for (int32_t c = 0; c < n_lookup; ++c) {
            C[(size_t)c] = (int32_t)acc;
            acc += cnt[(size_t)c];
        }
        C[(size_t)n_lookup] = (int32_t)length;
    }

    // --- suffix array
    std::vector<int32_t> sa_store;
    const int32_t* SA = sa_in;
    if (!SA && !parts) {
        sa_store.resize((size_t)length);
        suffix_array<uint16_t>(mapped.data(), sa_store.data(), (int32_t)length, n_codes);
        SA = sa_store.data();
    }
    double t1 = now_s();

    // --- sampling (:343-370)
    const int bw = min_bits((uint64_t)length);
    IntVec suffixes(length / sample_rate + 1, bw);
    IntVec positions;
    if (enable_extract) positions.init(length / sample_rate + 2, bw);
    BitString sampled;
    sampled.resize((uint64_t)length);
    if (parts) {
        if (parts->n_suffixes != (length - 1) / sample_rate + 1 || parts->n_suffixes > suffixes.length) {
            g_err = "sampled-row count does not match the text length";
            return -1;
        }
        for (int64_t k = 0; k < parts->n_suffixes; ++k) suffixes.set(k, (uint64_t)(uint32_t)parts->suffixes[k]);
        if (enable_extract) {
            if (!parts->positions) {
                g_err = "inverse-SA samples missing";
                return -1;
            }
            for (int64_t k = 0; k < positions.length; ++k) positions.set(k, (uint64_t)(uint32_t)parts->positions[k]);
        }
        sampled.from_words32(parts->mask_words, (uint64_t)length);
    } else {
        int64_t k = 0;
        for (int64_t i = 0; i < length; ++i) {
            int32_t p = SA[i];
            if (p % sample_rate == 0) {
                suffixes.set(k++, (uint64_t)p);
                sampled.set1((uint64_t)i);
                if (enable_extract) positions.set(p / sample_rate, (uint64_t)i);
            }
        }
        if (enable_extract) positions.set((length - 1) / sample_rate + 1, positions.get(0));
    }
    RrrEnc sampled_rrr;
    sampled_rrr.encode(sampled, sample_rate);
    sampled = BitString();

    // --- BWT (:374-394)
    std::vector<uint16_t> bwt((size_t)length);
    if (parts) {
        memcpy(bwt.data(), parts->bwt, (size_t)length * 2);
    } else {
        for (int64_t i = 0; i < length; ++i) {
            int32_t p = SA[i];
            bwt[(size_t)i] = p == 0 ? mapped[(size_t)(length - 1)] : mapped[(size_t)(p - 1)];
        }
    }
    sa_store = std::vector<int32_t>();
    mapped = std::vector<uint16_t>();
    double t2 = now_s();

    // --- wavelet structure (:173)
    int32_t wsigma = 0;
    for (int64_t i = 0; i < length; ++i) wsigma = std::max<int32_t>(wsigma, bwt[(size_t)i]);
    wsigma += 1;
    WfbbEnc w;
    wfbb_encode(bwt.data(), length, wsigma, sample_rate, threads, w);
    bwt = std::vector<uint16_t>();
    double t3 = now_s();

    // --- serialize (:948-975)
    JavaSink s(framed != 0);
    s.u8(0);
    s.i32(sample_rate);
    s.u8(enable_extract ? 1 : 0);
    s.i32(bw);
    s.i32(enable_extract ? bw : 0);
    s.i32((int32_t)length);
    std::vector<int32_t> keys = hashmap_key_order(insertion_keys);
    s.i32((int32_t)keys.size());
    for (int32_t k : keys) {
        s.i32(k);
        s.i16(code_of[(size_t)k]);
    }
    s.i32((int32_t)C.size());
    for (int32_t v : C) s.i32(v);
    s.i32((int32_t)lookup.size());
    for (int32_t v : lookup) s.i32(v);
    write_intvec(s, suffixes);
    if (enable_extract) write_intvec(s, positions);
    write_rrr(s, sampled_rrr);
    write_wfbb(s, w);
    s.finish();
    out_bytes.swap(s.out);
    double t4 = now_s();
    if (verbose)
        fprintf(stderr, "[fmhost] n=%lld sigma=%d: map+sa %.2fs, sample+bwt %.2fs, wfbb %.2fs, write %.2fs, %zu bytes\n",
                (long long)n, n_codes, t1 - t0, t2 - t1, t3 - t2, t4 - t3, out_bytes.size());
    return 0;
}

// splitmix64
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    inline uint64_t below(uint64_t n) { return next() % n; }
};

}  // namespace

extern "C" {

const char* fmhost_last_error(void) { return g_err.c_str(); }

// Builds a serialized FmIndex (reference layout).  *out is malloc'd; free with fmhost_free.
int fmhost_build(const uint16_t* text, int64_t n, int32_t sample_rate, int32_t enable_extract,
                 int32_t framed, int32_t threads, int32_t verbose, uint8_t** out, uint64_t* out_len) {
    std::vector<uint8_t> bytes;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    int rc = build_index(text, n, sample_rate, enable_extract, framed, threads, nullptr, verbose, bytes);
    if (rc) return rc;
    *out = (uint8_t*)malloc(bytes.size());
    if (!*out) {
        g_err = "out of memory";
        return -2;
    }
    memcpy(*out, bytes.data(), bytes.size());
    *out_len = bytes.size();
    return 0;
}

// Same, with a caller-supplied suffix array of (mapped text + sentinel) — used when the suffix
// array was produced on the GPU.
int fmhost_build_with_sa(const uint16_t* text, int64_t n, const int32_t* sa, int32_t sample_rate,
                         int32_t enable_extract, int32_t framed, int32_t threads, int32_t verbose,
                         uint8_t** out, uint64_t* out_len) {
    std::vector<uint8_t> bytes;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    int rc = build_index(text, n, sample_rate, enable_extract, framed, threads, sa, verbose, bytes);
    if (rc) return rc;
    *out = (uint8_t*)malloc(bytes.size());
    if (!*out) {
        g_err = "out of memory";
        return -2;
    }
    memcpy(*out, bytes.data(), bytes.size());
    *out_len = bytes.size();
    return 0;
}

// Same, from the pieces the device stage produced (fmgpu_build_bwt_samples_device, include/fmgpu.h): no suffix array crosses
// to the host.  `text` is still needed for the alphabet map and the cumulative counts.
int fmhost_build_with_parts(const uint16_t* text, int64_t n, const uint16_t* bwt, const uint32_t* mask_words, const int32_t* suffixes,
                            int64_t n_suffixes, const int32_t* positions, int32_t sample_rate, int32_t enable_extract, int32_t framed,
                            int32_t threads, int32_t verbose, uint8_t** out, uint64_t* out_len) {
    std::vector<uint8_t> bytes;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    BuildParts parts;
    parts.bwt = bwt;
    parts.mask_words = mask_words;
    parts.suffixes = suffixes;
    parts.n_suffixes = n_suffixes;
    parts.positions = positions;
    int rc = build_index(text, n, sample_rate, enable_extract, framed, threads, nullptr, verbose, bytes, &parts);
    if (rc) return rc;
    *out = (uint8_t*)malloc(bytes.size());
    if (!*out) {
        g_err = "out of memory";
        return -2;
    }
    memcpy(*out, bytes.data(), bytes.size());
    *out_len = bytes.size();
    return 0;
}

void fmhost_free(void* p) { free(p); }

static int emit(std::vector<uint8_t>& bytes, uint8_t** out, uint64_t* out_len) {
    *out = (uint8_t*)malloc(bytes.size() ? bytes.size() : 1);
    if (!*out) {
        g_err = "out of memory";
        return -2;
    }
    memcpy(*out, bytes.data(), bytes.size());
    *out_len = bytes.size();
    return 0;
}

// Stand-alone serialized WaveletFixedBlockBoosting over `syms` (values used as symbols directly, like
// new WaveletFixedBlockBoosting(short[] text, int samplingRate), WaveletFixedBlockBoosting.java:130-154).
int fmhost_build_wfbb(const uint16_t* syms, int64_t n, int32_t rrr_rate, int32_t framed, int32_t threads, uint8_t** out,
                      uint64_t* out_len) {
    if (n < 1) {
        g_err = "Input length must be > 0";
        return -1;
    }
    int32_t sigma = 0;
    for (int64_t i = 0; i < n; ++i) sigma = std::max<int32_t>(sigma, syms[i]);
    sigma += 1;
    WfbbEnc w;
    wfbb_encode(syms, n, sigma, rrr_rate, threads > 0 ? threads : 1, w);
    JavaSink s(framed != 0);
    write_wfbb(s, w);
    s.finish();
    return emit(s.out, out, out_len);
}

// Stand-alone serialized RrrVector over `nbits` LSB-first bits (new RrrVector(BitVector, sampleSize), RrrVector.java:225-286).
int fmhost_build_rrr(const uint64_t* words, int64_t nbits, int32_t sample_size, int32_t framed, uint8_t** out, uint64_t* out_len) {
    BitString bv;
    bv.resize((uint64_t)nbits);
    for (int64_t i = 0; i < (nbits + 63) / 64; ++i) bv.w[(size_t)i] = words[i];
    RrrEnc r;
    r.encode(bv, sample_size);
    JavaSink s(framed != 0);
    write_rrr(s, r);
    s.finish();
    return emit(s.out, out, out_len);
}

// First-appearance code mapping of a text (sentinel appended): what the suffix sorter must sort.
// codes_out has n+1 entries.  Returns the number of codes (alphabet incl. sentinel).
int32_t fmhost_map_text(const uint16_t* text, int64_t n, uint16_t* codes_out) {
    std::vector<int32_t> code_of(65536, -1);
    bool has_nul = false;
    for (int64_t i = 0; i < n; ++i)
        if (text[i] == 0) {
            has_nul = true;
            break;
        }
    int32_t next_code = has_nul ? 1 : 0;
    code_of[0] = next_code++;
    for (int64_t i = 0; i < n; ++i) {
        uint16_t c = text[i];
        if (code_of[c] < 0) code_of[c] = next_code++;
        codes_out[i] = (uint16_t)code_of[c];
    }
    codes_out[n] = 0;
    return next_code;
}

// Suffix array of codes[0..len) (last symbol unique smallest) — exposed for tests.
int fmhost_suffix_array(const uint16_t* codes, int32_t len, int32_t sigma, int32_t* sa_out) {
    suffix_array<uint16_t>(codes, sa_out, len, sigma);
    return 0;
}

// Synthetic log-like ASCII text: lines
//   "<yymmdd> <hhmmss> <pid> <LEVEL> <component>: <template with variable fields>\n"
// deterministic in `seed`; exactly n UTF-16 units are written (the last line is truncated).
void fmhost_gen_log_text(uint16_t* out, int64_t n, uint64_t seed) {
    static const char* levels[4] = {"INFO", "WARN", "ERROR", "DEBUG"};
    static const char* comps[] = {
        "dfs.DataNode$PacketResponder", "dfs.FSNamesystem", "dfs.DataNode$DataXceiver", "dfs.DataBlockScanner",
        "dfs.DataNode", "dfs.FSDataset", "ipc.Server", "mapred.JobTracker", "mapred.TaskTracker", "net.NetworkTopology",
        "security.UserGroupInformation", "http.HttpServer", "util.GSet", "namenode.FSEditLog", "namenode.LeaseManager",
        "balancer.Balancer", "blockmanagement.BlockManager", "datanode.BlockReceiver", "datanode.BlockSender",
        "metrics.MetricsSystemImpl", "yarn.ResourceManager", "yarn.NodeManager", "yarn.ContainerLauncher",
        "zk.ClientCnxn", "zk.QuorumPeer", "kafka.LogManager", "kafka.ReplicaFetcher", "db.ConnectionPool", "db.QueryPlanner",
        "cache.Evictor", "auth.TokenService", "auth.SessionStore", "rpc.Dispatcher", "rpc.RetryPolicy", "sched.Scheduler",
        "sched.Preemptor", "store.Compactor", "store.WAL", "gc.Monitor", "io.Throttler"};
    static const char* verbs[] = {"Receiving", "Received", "Served", "Deleting", "Verification succeeded for", "Starting",
                                  "Stopping", "Replicating", "Scheduling", "Allocated", "Released", "Committed",
                                  "Rolled back", "Opened", "Closed", "Flushed", "Compacted", "Evicted", "Renewed", "Rejected"};
    static const char* objs[] = {"block", "packet", "lease", "container", "task attempt", "segment", "connection",
                                 "session", "token", "partition"};
    static const char* tails[] = {"terminating", "of size", "from", "to", "src:", "dest:", "took", "ms", "retries",
                                  "is added to invalidSet of", "for user", "with status OK", "on queue", "at offset",
                                  "ack seqno", "bytes", "because quota exceeded", "after timeout", "in safe mode", "done"};
    const int NC = sizeof(comps) / sizeof(comps[0]), NV = sizeof(verbs) / sizeof(verbs[0]);
    const int NO = sizeof(objs) / sizeof(objs[0]), NT = sizeof(tails) / sizeof(tails[0]);
    Rng r(seed);
    int64_t pos = 0;
    uint64_t clock = 81109ULL * 86400ULL + 73000ULL;
    char line[512];
    while (pos < n) {
        clock += r.below(3);
        unsigned day = (unsigned)(clock / 86400ULL), sec = (unsigned)(clock % 86400ULL);
        unsigned lvl_roll = (unsigned)r.below(100);
        int lvl = lvl_roll < 80 ? 0 : lvl_roll < 92 ? 1 : lvl_roll < 97 ? 2 : 3;
        int tpl = (int)r.below(200);  // ~200 templates = (verb, object, tail, shape) tuples
        Rng tr(0xC0FFEEULL + (uint64_t)tpl);
        int comp = (int)tr.below(NC), verb = (int)tr.below(NV), obj = (int)tr.below(NO);
        int tail1 = (int)tr.below(NT), tail2 = (int)tr.below(NT), nfields = 1 + (int)tr.below(4);
        int len = snprintf(line, sizeof line, "%06u %02u%02u%02u %u %s %s: %s %s", day % 1000000u, sec / 3600, (sec / 60) % 60,
                           sec % 60, (unsigned)(1 + r.below(r.below(2) ? 99999 : 999)), levels[lvl], comps[comp],
                           verbs[verb], objs[obj]);
        for (int f = 0; f < nfields && len < 400; ++f) {
            int kind = (int)tr.below(3);
            if (kind == 0)
                len += snprintf(line + len, sizeof line - len, " blk_%lld", (long long)(r.next() >> 1) * (r.below(2) ? 1 : -1));
            else if (kind == 1)
                len += snprintf(line + len, sizeof line - len, " /10.%u.%u.%u:%u", (unsigned)r.below(256), (unsigned)r.below(256),
                                (unsigned)r.below(256), (unsigned)(1024 + r.below(60000)));
            else
                len += snprintf(line + len, sizeof line - len, " %llu", (unsigned long long)r.below(100000000ULL));
            if (f == 0) len += snprintf(line + len, sizeof line - len, " %s", tails[tail1]);
        }
        len += snprintf(line + len, sizeof line - len, " %s\n", tails[tail2]);
        for (int i = 0; i < len && pos < n; ++i) out[pos++] = (uint16_t)(unsigned char)line[i];
    }
}

// Query workload of the reference's JMH state (jmh/.../fm/FmIndexThroughputState.java:76-83):
// substrings of the text, uniform start in [0, n-max_len), uniform length in [min_len, max_len].
// Writes pattern chars concatenated into chars_out (capacity n_pat*max_len) and n_pat+1 offsets.
void fmhost_gen_patterns(const uint16_t* text, int64_t n, uint32_t n_pat, int32_t min_len, int32_t max_len,
                         uint64_t seed, uint16_t* chars_out, uint64_t* off_out) {
    Rng r(seed);
    uint64_t o = 0;
    for (uint32_t i = 0; i < n_pat; ++i) {
        int64_t start = (int64_t)r.below((uint64_t)(n - max_len));
        int32_t len = min_len + (int32_t)r.below((uint64_t)(max_len - min_len + 1));
        off_out[i] = o;
        memcpy(chars_out + o, text + start, sizeof(uint16_t) * (size_t)len);
        o += (uint64_t)len;
    }
    off_out[n_pat] = o;
}

}  // extern "C"
