// libfmgpu — C ABI (include/fmgpu.h) over the sm_100a kernels.  No CPU fallback: every entry point
// needs a CUDA device and fails with FMGPU_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fmgpu.h"
#include "flatten.hpp"
#include "jstream.hpp"
#include "kernels.cuh"
#include "kernels_lf.cuh"
#include "kernels_locate.cuh"
#include "kernels_utf8.cuh"
#include "kernels_wavelet.cuh"
#include "kernels_build.cuh"
#include "layout.h"

using namespace fmgpu;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(FMGPU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = prev == dev || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

enum { CTRL_QUEUE = 0, CTRL_QUEUE2 = 1, CTRL_STATS = 2 /* u64 x 9 at word 2.. */, CTRL_WORDS = 32 };

}  // namespace

enum { KIND_FM = 0, KIND_WAVELET = 1, KIND_RRR = 2 };

struct fmgpu_index {
    int device = 0;
    int kind = KIND_FM;  // what the handle was loaded from: an FmIndex stream, a bare WaveletFixedBlockBoosting, a bare RrrVector
    DevIndex dev{};
    std::vector<void*> allocs;
    uint64_t layout_bytes[8] = {0};
    uint64_t total_bytes = 0;
    int32_t alphabet_length = 0;
    int sm_count = 0;
    int count_ctas = 0, locate_ctas = 0, extract_ctas = 0, eub_ctas = 0;
    size_t tables_smem = 0;
    std::mutex mu;  // batch calls on one handle are serialised (v0)
    cudaStream_t stream = nullptr, copy_stream = nullptr, down_stream = nullptr;
    static constexpr int PIPE_SLOTS = 8;
    cudaEvent_t pipe_in[PIPE_SLOTS] = {nullptr}, pipe_out[PIPE_SLOTS] = {nullptr};
    Scratch codes, pats, ctrl, ranges, in_a, in_b, out_a, out_b, out_c, tmp_a, tmp_b, order, bins, u8conv;
    // The host-pointer count call runs its chunks on COUNT_CTX compute streams round-robin, each with its own scratch set,
    // so that the kernel of chunk k+1 fills the SMs as the longest patterns of chunk k drain (context 0 = the members above).
    static constexpr int COUNT_CTX = 4;
    struct CountCtx {
        Scratch pats, ctrl, order, bins;
    } cctx[COUNT_CTX - 1];
    cudaStream_t cstream[COUNT_CTX - 1] = {nullptr};
    // High-priority streams for the pre-pass kernels of a chunk (descriptors + length sort): they must not queue behind the
    // backward-search CTAs of earlier chunks, or the next search launch is late and the SMs drain.
    cudaStream_t pstream[COUNT_CTX] = {nullptr};
    cudaEvent_t pre_done[PIPE_SLOTS] = {nullptr};
    uint64_t last_launches = 0;
    bool stats_valid = false;
    bool count_stats = false;  // fmgpu_set_stats: backward-search kernel with work counters
    bool use_kmer = true;      // fmgpu_set_start_table: patterns start from the q-gram start table when the index has one
    uint32_t stats_ctx_mask = 1;  // compute contexts whose counters belong to the most recent call
    // optional per-call timing of the dominant kernel (bench.py's roofline): ring of event pairs
    static constexpr int TIMING_SLOTS = 64;
    bool timing = false;
    cudaEvent_t ev0[TIMING_SLOTS] = {nullptr}, ev1[TIMING_SLOTS] = {nullptr};
    uint64_t timed_calls = 0;
};

namespace {

template <typename T>
int upload(fmgpu_index* ix, const std::vector<T>& v, const T** dptr, uint64_t* bytes_acc) {
    void* p = nullptr;
    const size_t n = v.size() * sizeof(T);
    CU(cudaMalloc(&p, n ? n : 32));
    ix->allocs.push_back(p);
    if (n) CU(cudaMemcpy(p, v.data(), n, cudaMemcpyHostToDevice));
    *dptr = reinterpret_cast<const T*>(p);
    ix->total_bytes += n;
    if (bytes_acc) *bytes_acc += n;
    return 0;
}

int grid_for(const void* kernel, int sm_count, size_t smem, int* out) {
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, CTA_THREADS, smem));
    if (per_sm < 1) return fail(FMGPU_ERR_CUDA, "kernel does not fit on an SM");
    *out = per_sm * sm_count;
    return 0;
}

int prepass_grid(uint64_t items, int sm_count) {
    uint64_t g = (items + 255) / 256;
    const uint64_t cap = (uint64_t)sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// UTF-8 byte patterns: the pre-pass decodes d_bytes[pat_off[i], pat_off[i+1]) into d_chars at the same offsets (kernels_utf8.cuh)
struct Utf8Src {
    const uint8_t* d_bytes;
    uint16_t* d_chars;
    int32_t* d_conv_status;  // per pattern of this call: status of the conversion (0, 9, 10) ...
    int32_t* d_conv_value;   // ... and the offending code point
};

// Backward search over n_pat patterns on stream `st`.  `first_of_call` resets the work counters; later
// chunks of the same call only re-arm the work queue.
int count_on_stream(fmgpu_index* ix, const uint16_t* d_chars, const uint64_t* d_pat_off, uint64_t total_chars, uint32_t n_pat,
                    int32_t* d_counts, int32_t* d_status, uint32_t* d_ranges, cudaStream_t st, bool first_of_call = true, int ctx = 0,
                    const Utf8Src* u8 = nullptr, cudaStream_t pre = nullptr, cudaEvent_t pre_ev = nullptr, int threads = CTA_THREADS) {
    // `pre` (optional): stream for the pre-pass kernels, joined into `st` through pre_ev before the search kernel
    (void)total_chars;
    if (ix->kind != KIND_FM) return fail(FMGPU_ERR_UNSUPPORTED, "the handle holds no FmIndex (loaded from a bare wavelet / RRR stream)");
    if (u8) d_chars = u8->d_chars;
    Scratch& s_pats = ctx ? ix->cctx[ctx - 1].pats : ix->pats;
    Scratch& s_ctrl = ctx ? ix->cctx[ctx - 1].ctrl : ix->ctrl;
    Scratch& s_order = ctx ? ix->cctx[ctx - 1].order : ix->order;
    Scratch& s_bins = ctx ? ix->cctx[ctx - 1].bins : ix->bins;
    CU(s_pats.reserve(((size_t)n_pat + 2) * sizeof(PatDesc)));
    CU(s_ctrl.reserve(CTRL_WORDS * 4));
    CU(s_order.reserve((size_t)n_pat * 4 + 64));
    CU(s_bins.reserve(LEN_BINS * 4));
    if (!pre) pre = st;
    if (first_of_call) {
        CU(cudaMemsetAsync(s_ctrl.p, 0, CTRL_WORDS * 4, pre));
        if (ctx == 0) {
            ix->last_launches = 0;
            ix->stats_valid = true;
            ix->stats_ctx_mask = 1;
        } else {
            ix->stats_ctx_mask |= 1u << ctx;
        }
    } else {
        CU(cudaMemsetAsync(s_ctrl.p, 0, 8, pre));  // the two queue heads
    }
    if (n_pat == 0) {
        if (pre != st) {
            CU(cudaEventRecord(pre_ev, pre));
            CU(cudaStreamWaitEvent(st, pre_ev, 0));
        }
        return 0;
    }
    CU(cudaMemsetAsync(s_bins.p, 0, LEN_BINS * 4, pre));
    unsigned int* ctrl = (unsigned int*)s_ctrl.p;
    // descriptors + length histogram, then a counting sort by length so that a warp's 32 patterns run in lockstep
    const int pre_grid = prepass_grid(n_pat, ix->sm_count);
    const uint32_t kq = ix->use_kmer ? ix->dev.kmer_q : 0u;  // 0: every pattern starts from its last char
    if (u8)
        k_prepass_utf8<<<pre_grid, 256, 0, pre>>>(u8->d_bytes, d_pat_off, n_pat, ix->dev.char2code, u8->d_chars, (PatDesc*)s_pats.p,
                                                 (uint32_t*)s_bins.p, u8->d_conv_status, u8->d_conv_value, kq, ix->dev.kmer_stride, ix->dev.sigma);
    else
        k_prepass<<<pre_grid, 256, 0, pre>>>(d_chars, d_pat_off, n_pat, ix->dev.char2code, (PatDesc*)s_pats.p, (uint32_t*)s_bins.p, kq,
                                             ix->dev.kmer_stride, ix->dev.sigma);
    k_len_scan<<<1, SCAN_THREADS, 0, pre>>>((uint32_t*)s_bins.p);
    const int sc_grid = prepass_grid(((uint64_t)n_pat + SCATTER_PER_THREAD - 1) / SCATTER_PER_THREAD, ix->sm_count);
    k_len_scatter<<<sc_grid, 256, 0, pre>>>((const PatDesc*)s_pats.p, n_pat, (uint32_t*)s_bins.p, (uint32_t*)s_order.p);
    if (pre != st) {
        CU(cudaEventRecord(pre_ev, pre));
        CU(cudaStreamWaitEvent(st, pre_ev, 0));
    }
    const int slot = (int)(ix->timed_calls % fmgpu_index::TIMING_SLOTS);
    if (ix->timing) CU(cudaEventRecord(ix->ev0[slot], st));
    int grid = ix->count_ctas;
    const int need = (int)(((uint64_t)n_pat + threads - 1) / threads);
    if (need < grid) grid = need;
    if (ix->count_stats)
        k_count<true><<<grid, threads, ix->tables_smem, st>>>(ix->dev, d_chars, (const PatDesc*)s_pats.p, (const uint32_t*)s_order.p, n_pat,
                                                                  d_counts, d_status, d_ranges, ctrl + CTRL_QUEUE,
                                                                  (unsigned long long*)(ctrl + CTRL_STATS));
    else
        k_count<false><<<grid, threads, ix->tables_smem, st>>>(ix->dev, d_chars, (const PatDesc*)s_pats.p, (const uint32_t*)s_order.p, n_pat,
                                                                   d_counts, d_status, d_ranges, ctrl + CTRL_QUEUE,
                                                                   (unsigned long long*)(ctrl + CTRL_STATS));
    if (ix->timing) {
        CU(cudaEventRecord(ix->ev1[slot], st));
        ix->timed_calls++;
    }
    ix->last_launches += 4;
    if (u8) {  // a pattern whose conversion throws never reaches the search in the reference: its status wins
        k_utf8_merge<<<(n_pat + 255) / 256, 256, 0, st>>>(u8->d_conv_status, u8->d_conv_value, n_pat, d_counts, d_status, d_ranges ? 0 : 1);
        ix->last_launches += 1;
    }
    CU(cudaGetLastError());
    return 0;
}

constexpr int START_TABLE_LOG2 = 25;  // up to 32 M entries = 256 MB (q = 4 for a 70-symbol log alphabet: 24 M entries)

// q-gram start table (layout.h): the search kernel itself computes, for every q-gram of alphabet codes, the SA range after its
// q chars; q = the largest value with sigma^q entries <= 2^21 (q >= 2).  Built once per load (~1 ms of GPU time).
int build_start_table(fmgpu_index* ix, const std::vector<uint16_t>& code2char, const std::vector<uint16_t>& char2code) {
    if (const char* e = getenv("FMGPU_START_TABLE"))
        if (atoi(e) == 0) return 0;
    const uint64_t S = ix->dev.sigma;
    int log2_max = START_TABLE_LOG2;  // FMGPU_START_TABLE_LOG2: largest table, in log2 entries of 8 bytes
    if (const char* e = getenv("FMGPU_START_TABLE_LOG2")) log2_max = atoi(e) >= 2 && atoi(e) <= 27 ? atoi(e) : log2_max;
    {  // no more than ~4 entries per text position: a small index does not get a table larger than itself
        int lb = 2;
        while ((1ull << lb) < (uint64_t)ix->dev.length * 4 && lb < 27) ++lb;
        if (lb < log2_max) log2_max = lb;
    }
    if (S < 2 || S * S > (1ull << log2_max)) return 0;
    uint32_t q = 2;
    while (q < 8 && [&] { uint64_t n = 1; for (uint32_t k = 0; k <= q; ++k) n *= S; return n; }() <= (1ull << log2_max)) ++q;
    uint64_t n_entries = 1;
    for (uint32_t k = 0; k < q; ++k) n_entries *= S;
    // every q-gram of codes 1 .. S-1 that chars can spell; pattern text order = first-consumed char last
    std::vector<uint32_t> idx_of;
    std::vector<uint16_t> chars;
    std::vector<uint32_t> codes(q);
    for (uint64_t idx = 0; idx < n_entries; ++idx) {
        uint64_t v = idx;
        bool ok = true;
        for (uint32_t k = 0; k < q; ++k) {  // codes[0] = last consumed ... codes[q-1] = the q-gram's last char
            codes[k] = (uint32_t)(v % S);
            v /= S;
            if (codes[k] == 0 || codes[k] >= code2char.size() || char2code[code2char[codes[k]]] != codes[k]) ok = false;
        }
        if (!ok) continue;
        idx_of.push_back((uint32_t)idx);
        for (uint32_t k = 0; k < q; ++k) chars.push_back(code2char[codes[k]]);
    }
    const uint32_t n = (uint32_t)idx_of.size();
    std::vector<U32x2> table((size_t)n_entries, U32x2{0xffffffffu, 0u});
    if (n) {
        std::vector<uint64_t> off((size_t)n + 1);
        for (uint32_t i = 0; i <= n; ++i) off[i] = (uint64_t)i * q;
        uint16_t* d_chars = nullptr;
        uint64_t* d_off = nullptr;
        int32_t *d_counts = nullptr, *d_status = nullptr;
        uint32_t* d_ranges = nullptr;
        cudaStream_t st = ix->stream;
        CU(cudaMalloc((void**)&d_chars, chars.size() * 2));
        CU(cudaMalloc((void**)&d_off, off.size() * 8));
        CU(cudaMalloc((void**)&d_counts, (size_t)n * 4));
        CU(cudaMalloc((void**)&d_status, (size_t)n * 4));
        CU(cudaMalloc((void**)&d_ranges, (size_t)n * 8));
        CU(cudaMemcpyAsync(d_chars, chars.data(), chars.size() * 2, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
        const bool was = ix->use_kmer;
        ix->use_kmer = false;
        int rc = count_on_stream(ix, d_chars, d_off, chars.size(), n, d_counts, d_status, d_ranges, st);
        ix->use_kmer = was;
        std::vector<int32_t> status(n);
        std::vector<uint32_t> ranges((size_t)n * 2);
        if (!rc) {
            CU(cudaMemcpyAsync(status.data(), d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(ranges.data(), d_ranges, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        for (void* p : {(void*)d_chars, (void*)d_off, (void*)d_counts, (void*)d_status, (void*)d_ranges}) cudaFree(p);
        if (n > (1u << 22))  // a large table was built with scratch buffers no query batch is likely to need: give them back
            for (Scratch* sc : {&ix->pats, &ix->order}) sc->release();
        if (rc) return rc;
        for (uint32_t i = 0; i < n; ++i)
            if (status[i] == 0) table[idx_of[i]] = U32x2{ranges[2 * (size_t)i], ranges[2 * (size_t)i + 1]};
    }
    int rc = upload(ix, table, &ix->dev.kmer, nullptr);
    if (rc) return rc;
    ix->dev.kmer_q = q;
    ix->dev.kmer_stride = (uint32_t)S;
    ix->stats_valid = false;
    return 0;
}

}  // namespace

#include "api_lf.inc"
#include "api_wavelet.inc"
#include "api_build.inc"

extern "C" {

const char* fmgpu_last_error(void) { return g_err.c_str(); }
const char* fmgpu_version(void) { return "fmgpu 0.1 (sm_100a)"; }

}  // extern "C"

namespace {
// Parses a serialized FmIndex / WaveletFixedBlockBoosting / RrrVector (Java stream layout), re-lays it out and uploads it.
int load_common(const uint8_t* buf, size_t len, const fmgpu_opts* opts, fmgpu_index** out, int kind) {
    if (!buf || !out) return fail(FMGPU_ERR_ARG, "null argument");
    *out = nullptr;
    int dev = opts ? opts->device : -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(FMGPU_ERR_CUDA, "no CUDA device available (libfmgpu has no CPU fallback)");
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(FMGPU_ERR_ARG, "device ordinal %d out of range", dev);
    int threads = opts ? opts->host_threads : 0;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;

    fmgpu_host::FlatIndex F;
    try {
        fmgpu_host::JavaIn in(buf, len);
        if (kind == KIND_FM) {
            fmgpu_host::FmStream fm;
            fm.read(in);
            fmgpu_host::flatten(fm, threads, F);
        } else if (kind == KIND_WAVELET) {  // WaveletFixedBlockBoosting.read (wavelet/WaveletFixedBlockBoosting.java:286-322)
            fmgpu_host::FmStream fm;
            fm.wf.read(in);
            fm.length = (int32_t)fm.wf.size;
            fm.sample_rate = 1;
            fmgpu_host::flatten(fm, threads, F, true);
        } else {  // RrrVector.read (bitsequence/RrrVector.java:448-469)
            fmgpu_host::RrrStream r;
            r.read(in);
            fmgpu_host::flatten_rrr(r, F);
        }
    } catch (const fmgpu_host::FormatError& e) {
        return fail(FMGPU_ERR_FORMAT, "%s", e.what());
    } catch (const std::bad_alloc&) {
        return fail(FMGPU_ERR_FORMAT, "out of host memory while re-laying out the index");
    } catch (const std::exception& e) {
        return fail(FMGPU_ERR_FORMAT, "%s", e.what());
    }

    DeviceGuard g(dev);
    if (!g.ok) return fail(FMGPU_ERR_CUDA, "cannot select device %d", dev);
    fmgpu_index* ix = new fmgpu_index();
    ix->device = dev;
    ix->kind = kind;
    ix->dev = F.meta;
    ix->alphabet_length = F.alphabet_length;
    int rc = 0;
    auto up = [&](int r) {
        if (!rc) rc = r;
    };
    up(upload(ix, F.C, &ix->dev.C, nullptr));
    up(upload(ix, F.char2code, &ix->dev.char2code, nullptr));
    up(upload(ix, F.code2char, &ix->dev.code2char, nullptr));
    up(upload(ix, F.sb, &ix->dev.sb, nullptr));
    up(upload(ix, F.sbroot, &ix->dev.sbroot, nullptr));
    up(upload(ix, F.blkmap, &ix->dev.blkmap, nullptr));
    up(upload(ix, F.cells, &ix->dev.cells, &ix->layout_bytes[0]));
    up(upload(ix, F.sectors, &ix->dev.sectors, &ix->layout_bytes[1]));
    up(upload(ix, F.nodes, &ix->dev.nodes, &ix->layout_bytes[2]));
    up(upload(ix, F.blocks, &ix->dev.blocks, &ix->layout_bytes[3]));
    up(upload(ix, F.ovf, &ix->dev.ovf, &ix->layout_bytes[4]));
    up(upload(ix, F.sgroups, &ix->dev.sgroups, &ix->layout_bytes[5]));
    up(upload(ix, F.soffsets, &ix->dev.soffsets, &ix->layout_bytes[5]));
    {
        const fmgpu_host::RrrTables& RT = fmgpu_host::rrr_tables();
        std::vector<uint16_t> inv(RT.inverse, RT.inverse + 32768), cb(RT.class_base, RT.class_base + 16);
        up(upload(ix, inv, &ix->dev.rrr_inv, nullptr));
        up(upload(ix, cb, &ix->dev.rrr_cbase, nullptr));
    }
    up(upload(ix, F.sa, &ix->dev.sa, &ix->layout_bytes[6]));
    up(upload(ix, F.isa, &ix->dev.isa, &ix->layout_bytes[7]));
    if (!rc) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) rc = fail(FMGPU_ERR_CUDA, "cudaGetDeviceProperties failed");
        else ix->sm_count = prop.multiProcessorCount;
    }
    if (!rc) {
        ix->tables_smem = count_smem_bytes(ix->dev);
        // the attribute is per function, not per index: always the largest table set any index can stage
        if (cudaFuncSetAttribute((const void*)k_count<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COUNT_SMEM_MAX_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute((const void*)k_count<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COUNT_SMEM_MAX_BYTES) != cudaSuccess)
            rc = fail(FMGPU_ERR_CUDA, "k_count: cannot reserve %zu bytes of shared memory", ix->tables_smem);
        int g0 = 0, g1 = 0;
        if (!rc) rc = grid_for((const void*)k_count<false>, ix->sm_count, ix->tables_smem, &g0);
        if (!rc) rc = grid_for((const void*)k_count<true>, ix->sm_count, ix->tables_smem, &g1);
        ix->count_ctas = g0 < g1 ? g0 : g1;
    }
    if (!rc) rc = lf_setup(ix);
    if (!rc && (cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamCreateWithFlags(&ix->down_stream, cudaStreamNonBlocking) != cudaSuccess))
        rc = fail(FMGPU_ERR_CUDA, "cudaStreamCreate failed");
    for (int i = 0; !rc && i < fmgpu_index::COUNT_CTX - 1; ++i)
        if (cudaStreamCreateWithFlags(&ix->cstream[i], cudaStreamNonBlocking) != cudaSuccess) rc = fail(FMGPU_ERR_CUDA, "cudaStreamCreate failed");
    if (!rc) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lower = higher priority
        for (int i = 0; !rc && i < fmgpu_index::COUNT_CTX; ++i)
            if (cudaStreamCreateWithPriority(&ix->pstream[i], cudaStreamNonBlocking, prio_hi) != cudaSuccess)
                rc = fail(FMGPU_ERR_CUDA, "cudaStreamCreateWithPriority failed");
    }
    for (int i = 0; !rc && i < fmgpu_index::PIPE_SLOTS; ++i)
        if (cudaEventCreateWithFlags(&ix->pipe_in[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ix->pre_done[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ix->pipe_out[i], cudaEventDisableTiming) != cudaSuccess)
            rc = fail(FMGPU_ERR_CUDA, "cudaEventCreate failed");
    if (!rc) {
        // optional experiment knob: L2 fetch granularity for the random 32-byte record gathers
        const char* g = getenv("FMGPU_L2_FETCH");
        if (g && atoi(g) > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    }
    if (!rc && kind == KIND_FM) rc = build_start_table(ix, F.code2char, F.char2code);
    if (rc) {
        fmgpu_index_free(ix);
        return rc;
    }
    *out = ix;
    return 0;
}
}  // namespace

extern "C" {

int fmgpu_index_load_serialized(const uint8_t* buf, size_t len, const fmgpu_opts* opts, fmgpu_index** out) {
    return load_common(buf, len, opts, out, KIND_FM);
}
int fmgpu_wavelet_load_serialized(const uint8_t* buf, size_t len, const fmgpu_opts* opts, fmgpu_index** out) {
    return load_common(buf, len, opts, out, KIND_WAVELET);
}
int fmgpu_rrr_load_serialized(const uint8_t* buf, size_t len, const fmgpu_opts* opts, fmgpu_index** out) {
    return load_common(buf, len, opts, out, KIND_RRR);
}

void fmgpu_index_free(fmgpu_index* ix) {
    if (!ix) return;
    DeviceGuard g(ix->device);
    for (void* p : ix->allocs) cudaFree(p);
    for (Scratch* s : {&ix->codes, &ix->pats, &ix->ctrl, &ix->ranges, &ix->in_a, &ix->in_b, &ix->out_a, &ix->out_b, &ix->out_c,
                       &ix->tmp_a, &ix->tmp_b, &ix->order, &ix->bins, &ix->u8conv})
        s->release();
    for (int i = 0; i < fmgpu_index::COUNT_CTX - 1; ++i) {
        for (Scratch* sc : {&ix->cctx[i].pats, &ix->cctx[i].ctrl, &ix->cctx[i].order, &ix->cctx[i].bins}) sc->release();
        if (ix->cstream[i]) cudaStreamDestroy(ix->cstream[i]);
    }
    for (int i = 0; i < fmgpu_index::COUNT_CTX; ++i)
        if (ix->pstream[i]) cudaStreamDestroy(ix->pstream[i]);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    if (ix->copy_stream) cudaStreamDestroy(ix->copy_stream);
    if (ix->down_stream) cudaStreamDestroy(ix->down_stream);
    for (int i = 0; i < fmgpu_index::PIPE_SLOTS; ++i) {
        if (ix->pipe_in[i]) cudaEventDestroy(ix->pipe_in[i]);
        if (ix->pipe_out[i]) cudaEventDestroy(ix->pipe_out[i]);
        if (ix->pre_done[i]) cudaEventDestroy(ix->pre_done[i]);
    }
    for (int i = 0; i < fmgpu_index::TIMING_SLOTS; ++i) {
        if (ix->ev0[i]) cudaEventDestroy(ix->ev0[i]);
        if (ix->ev1[i]) cudaEventDestroy(ix->ev1[i]);
    }
    delete ix;
}

int32_t fmgpu_input_length(const fmgpu_index* ix) { return ix ? (int32_t)ix->dev.length : -1; }
int32_t fmgpu_alphabet_length(const fmgpu_index* ix) { return ix ? ix->alphabet_length : -1; }
int32_t fmgpu_sample_rate(const fmgpu_index* ix) { return ix ? (int32_t)ix->dev.sample_rate : -1; }
int32_t fmgpu_extract_enabled(const fmgpu_index* ix) { return ix ? (int32_t)ix->dev.extract_enabled : -1; }
int32_t fmgpu_device(const fmgpu_index* ix) { return ix ? ix->device : -1; }
uint64_t fmgpu_device_bytes(const fmgpu_index* ix) { return ix ? ix->total_bytes : 0; }
void fmgpu_layout_bytes(const fmgpu_index* ix, uint64_t out8[8]) {
    for (int i = 0; i < 8; ++i) out8[i] = ix ? ix->layout_bytes[i] : 0;
}

int fmgpu_count_batch_device(fmgpu_index* ix, const uint16_t* d_chars, const uint64_t* d_pat_off, uint64_t total_chars, uint32_t n_pat,
                             int32_t* d_counts_out, int32_t* d_status_out, void* cuda_stream) {
    if (!ix || !d_pat_off || !d_counts_out || (!d_chars && total_chars)) return fail(FMGPU_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (!g.ok) return fail(FMGPU_ERR_CUDA, "cannot select device %d", ix->device);
    return count_on_stream(ix, d_chars, d_pat_off, total_chars, n_pat, d_counts_out, d_status_out, nullptr, (cudaStream_t)cuda_stream);
}

namespace {
// Host-pointer count call for char[] patterns (unit 2) or UTF-8 byte patterns (unit 1).
int count_host(fmgpu_index* ix, const void* in, size_t unit, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts_out, int32_t* status_out) {
    if (!ix || !pat_off || !counts_out) return fail(FMGPU_ERR_ARG, "null argument");
    const uint64_t total = pat_off[n_pat];
    if (total && !in) return fail(FMGPU_ERR_ARG, "null argument");
    const bool utf8 = unit == 1;
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (!g.ok) return fail(FMGPU_ERR_CUDA, "cannot select device %d", ix->device);
    cudaStream_t st = ix->stream, cp = ix->copy_stream;
    CU(ix->in_a.reserve((size_t)total * 2 + 64));
    CU(ix->in_b.reserve(((size_t)n_pat + 1) * 8));
    CU(ix->out_a.reserve((size_t)n_pat * 4 + 64));
    CU(ix->out_b.reserve((size_t)n_pat * 4 + 64));
    if (utf8) {
        CU(ix->codes.reserve((size_t)total + 64));
        CU(ix->u8conv.reserve((size_t)n_pat * 8 + 64));
    }
    // The batch is cut into chunks: the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the kernels of chunk k
    // (copy stream, COUNT_CTX compute streams round-robin, download stream, events in between).  A launch lasts at least as
    // long as its longest pattern's dependent chain (~0.2 ms), so consecutive chunks run on different compute streams and the
    // next chunk's CTAs fill the SMs while the previous chunk's last warps drain.  Chunk chars land at their absolute offsets
    // in the device buffer, so pattern offsets need no rebasing.
    uint32_t min_chunk = 125000;
    if (const char* e = getenv("FMGPU_PIPE_CHUNK")) min_chunk = (uint32_t)atoi(e) > 0 ? (uint32_t)atoi(e) : min_chunk;
    int n_ctx = fmgpu_index::COUNT_CTX;
    if (const char* e = getenv("FMGPU_PIPE_STREAMS")) n_ctx = atoi(e) >= 1 && atoi(e) <= fmgpu_index::COUNT_CTX ? atoi(e) : n_ctx;
    uint32_t n_chunks = n_pat / min_chunk;
    if (n_chunks > (uint32_t)fmgpu_index::PIPE_SLOTS) n_chunks = fmgpu_index::PIPE_SLOTS;
    if (n_chunks < 1) n_chunks = 1;
    // search CTAs of the chunked call are a little smaller than CTA_THREADS so that two of them leave registers and thread
    // slots on the SM for the (small) pre-pass CTAs of the next chunks
    int pipe_threads = n_chunks > 1 ? PIPE_CTA_THREADS : CTA_THREADS;
    if (const char* e = getenv("FMGPU_PIPE_THREADS")) pipe_threads = atoi(e) >= 32 && atoi(e) <= CTA_THREADS ? (atoi(e) / 32) * 32 : pipe_threads;
    uint16_t* d_chars = (uint16_t*)ix->in_a.p;
    uint8_t* d_in = utf8 ? (uint8_t*)ix->codes.p : (uint8_t*)ix->in_a.p;
    uint64_t* d_off = (uint64_t*)ix->in_b.p;
    int32_t* d_counts = (int32_t*)ix->out_a.p;
    int32_t* d_status = (int32_t*)ix->out_b.p;
    const uint8_t* h_in = (const uint8_t*)in;
    // FMGPU_PIPE_TRACE=1: per-chunk timeline of the call on stderr (timing events; diagnostic only)
    const bool trace = getenv("FMGPU_PIPE_TRACE") != nullptr;
    cudaEvent_t t0 = nullptr, t_in[fmgpu_index::PIPE_SLOTS], t_k0[fmgpu_index::PIPE_SLOTS], t_k1[fmgpu_index::PIPE_SLOTS], t_out[fmgpu_index::PIPE_SLOTS];
    if (trace) {
        CU(cudaEventCreate(&t0));
        for (uint32_t k = 0; k < n_chunks; ++k) {
            CU(cudaEventCreate(&t_in[k]));
            CU(cudaEventCreate(&t_k0[k]));
            CU(cudaEventCreate(&t_k1[k]));
            CU(cudaEventCreate(&t_out[k]));
        }
        CU(cudaEventRecord(t0, cp));
    }
    for (uint32_t k = 0; k < n_chunks; ++k) {  // all uploads are queued first: they only depend on the host buffers
        const uint32_t lo = (uint32_t)((uint64_t)n_pat * k / n_chunks), hi = (uint32_t)((uint64_t)n_pat * (k + 1) / n_chunks);
        const uint64_t c0 = pat_off[lo], c1 = pat_off[hi];
        if (c1 > c0) CU(cudaMemcpyAsync(d_in + c0 * unit, h_in + c0 * unit, (size_t)(c1 - c0) * unit, cudaMemcpyHostToDevice, cp));
        CU(cudaMemcpyAsync(d_off + lo, pat_off + lo, ((size_t)(hi - lo) + 1) * 8, cudaMemcpyHostToDevice, cp));
        CU(cudaEventRecord(ix->pipe_in[k], cp));
        if (trace) CU(cudaEventRecord(t_in[k], cp));
    }
    for (uint32_t k = 0; k < n_chunks; ++k) {
        const uint32_t lo = (uint32_t)((uint64_t)n_pat * k / n_chunks), hi = (uint32_t)((uint64_t)n_pat * (k + 1) / n_chunks);
        const int ctx = (int)(k % (uint32_t)n_ctx);
        cudaStream_t cs = ctx ? ix->cstream[ctx - 1] : st;
        cudaStream_t ps = ix->pstream[ctx];
        // the pre-pass of chunk k runs on the context's high-priority stream as soon as the chunk has arrived and the previous
        // search of this context (chunk k - n_ctx, same scratch buffers) is done
        CU(cudaStreamWaitEvent(ps, ix->pipe_in[k], 0));
        if (k >= (uint32_t)n_ctx) CU(cudaStreamWaitEvent(ps, ix->pipe_out[k - n_ctx], 0));
        Utf8Src u8{d_in, d_chars, (int32_t*)ix->u8conv.p + lo, (int32_t*)ix->u8conv.p + n_pat + lo};
        if (trace) CU(cudaEventRecord(t_k0[k], ps));
        int rc = count_on_stream(ix, d_chars, d_off + lo, total, hi - lo, d_counts + lo, d_status + lo, nullptr, cs, k < (uint32_t)n_ctx, ctx,
                                 utf8 ? &u8 : nullptr, ps, ix->pre_done[k], pipe_threads);
        if (rc) return rc;
        CU(cudaEventRecord(ix->pipe_out[k], cs));
        if (trace) CU(cudaEventRecord(t_k1[k], cs));
        CU(cudaStreamWaitEvent(ix->down_stream, ix->pipe_out[k], 0));
        if (hi > lo) {
            CU(cudaMemcpyAsync(counts_out + lo, d_counts + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, ix->down_stream));
            if (status_out)
                CU(cudaMemcpyAsync(status_out + lo, d_status + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, ix->down_stream));
        }
        if (trace) CU(cudaEventRecord(t_out[k], ix->down_stream));
    }
    CU(cudaStreamSynchronize(ix->down_stream));
    CU(cudaStreamSynchronize(cp));
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < fmgpu_index::COUNT_CTX - 1; ++i) CU(cudaStreamSynchronize(ix->cstream[i]));
    for (int i = 0; i < fmgpu_index::COUNT_CTX; ++i) CU(cudaStreamSynchronize(ix->pstream[i]));
    if (trace) {
        for (uint32_t k = 0; k < n_chunks; ++k) {
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime(&a, t0, t_in[k]);
            cudaEventElapsedTime(&b, t0, t_k0[k]);
            cudaEventElapsedTime(&c, t0, t_k1[k]);
            cudaEventElapsedTime(&d, t0, t_out[k]);
            fprintf(stderr, "[fmgpu trace] chunk %u: h2d done %.3f ms, kernels %.3f .. %.3f ms, d2h done %.3f ms\n", k, a, b, c, d);
            cudaEventDestroy(t_in[k]);
            cudaEventDestroy(t_k0[k]);
            cudaEventDestroy(t_k1[k]);
            cudaEventDestroy(t_out[k]);
        }
        cudaEventDestroy(t0);
    }
    return 0;
}
}  // namespace

int fmgpu_count_batch(fmgpu_index* ix, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts_out,
                      int32_t* status_out) {
    return count_host(ix, chars, 2, pat_off, n_pat, counts_out, status_out);
}

int fmgpu_count_batch_utf8(fmgpu_index* ix, const uint8_t* bytes, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts_out,
                           int32_t* status_out) {
    return count_host(ix, bytes, 1, pat_off, n_pat, counts_out, status_out);
}

int fmgpu_count_batch_utf8_device(fmgpu_index* ix, const uint8_t* d_bytes, const uint64_t* d_pat_off, uint64_t total_bytes, uint32_t n_pat,
                                  int32_t* d_counts_out, int32_t* d_status_out, void* cuda_stream) {
    if (!ix || !d_pat_off || !d_counts_out || (!d_bytes && total_bytes)) return fail(FMGPU_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (!g.ok) return fail(FMGPU_ERR_CUDA, "cannot select device %d", ix->device);
    CU(ix->in_a.reserve((size_t)total_bytes * 2 + 64));
    CU(ix->u8conv.reserve((size_t)n_pat * 8 + 64));
    Utf8Src u8{d_bytes, (uint16_t*)ix->in_a.p, (int32_t*)ix->u8conv.p, (int32_t*)ix->u8conv.p + n_pat};
    return count_on_stream(ix, nullptr, d_pat_off, total_bytes, n_pat, d_counts_out, d_status_out, nullptr, (cudaStream_t)cuda_stream, true, 0, &u8);
}

int fmgpu_set_start_table(fmgpu_index* ix, int enable) {
    if (!ix) return fail(FMGPU_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->use_kmer = enable != 0;
    return 0;
}
int32_t fmgpu_start_table_q(const fmgpu_index* ix) { return ix ? (int32_t)ix->dev.kmer_q : -1; }

int fmgpu_set_stats(fmgpu_index* ix, int enable) {
    if (!ix) return fail(FMGPU_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->count_stats = enable != 0;
    return 0;
}

int fmgpu_set_timing(fmgpu_index* ix, int enable) {
    if (!ix) return fail(FMGPU_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (enable && !ix->ev0[0]) {
        for (int i = 0; i < fmgpu_index::TIMING_SLOTS; ++i) {
            CU(cudaEventCreate(&ix->ev0[i]));
            CU(cudaEventCreate(&ix->ev1[i]));
        }
    }
    ix->timing = enable != 0;
    ix->timed_calls = 0;
    return 0;
}

int fmgpu_search_kernel_ms(fmgpu_index* ix, uint32_t calls_back, float* ms_out) {
    if (!ix || !ms_out) return fail(FMGPU_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (!ix->ev0[0] || calls_back >= ix->timed_calls || calls_back >= (uint32_t)fmgpu_index::TIMING_SLOTS)
        return fail(FMGPU_ERR_ARG, "no timing recorded for that call");
    const int slot = (int)((ix->timed_calls - 1 - calls_back) % fmgpu_index::TIMING_SLOTS);
    CU(cudaEventSynchronize(ix->ev1[slot]));
    CU(cudaEventElapsedTime(ms_out, ix->ev0[slot], ix->ev1[slot]));
    return 0;
}

int fmgpu_last_stats(fmgpu_index* ix, uint64_t out8[8]) {
    uint64_t v[FMGPU_N_STATS];
    if (!out8) return fail(FMGPU_ERR_ARG, "null argument");
    const int rc = fmgpu_last_stats_ex(ix, v, FMGPU_N_STATS);
    if (!rc) memcpy(out8, v, 8 * sizeof(uint64_t));
    return rc;
}

int fmgpu_last_stats_ex(fmgpu_index* ix, uint64_t* out8, uint32_t n_out) {
    if (!ix || !out8 || n_out < FMGPU_N_STATS) return fail(FMGPU_ERR_ARG, "null argument or fewer than FMGPU_N_STATS slots");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    memset(out8, 0, FMGPU_N_STATS * sizeof(uint64_t));
    if (!ix->stats_valid || !ix->ctrl.p) return 0;
    CU(cudaDeviceSynchronize());
    for (int c = 0; c < fmgpu_index::COUNT_CTX; ++c) {
        if (!(ix->stats_ctx_mask & (1u << c))) continue;
        void* p = c ? ix->cctx[c - 1].ctrl.p : ix->ctrl.p;
        if (!p) continue;
        uint32_t words[CTRL_WORDS];
        CU(cudaMemcpy(words, p, sizeof words, cudaMemcpyDeviceToHost));
        uint64_t v[FMGPU_N_STATS];
        memcpy(v, words + CTRL_STATS, FMGPU_N_STATS * sizeof(uint64_t));
        for (int i = 0; i < FMGPU_N_STATS; ++i)
            if (i != 5) out8[i] += v[i];
    }
    out8[5] = ix->last_launches;
    return 0;
}

}  // extern "C"
