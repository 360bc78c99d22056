// libfmgpu — C ABI (include/fmgpu.h) over the sm_100a kernels.  No CPU fallback: every entry point
// needs a CUDA device and fails with FMGPU_ERR_CUDA otherwise.
//
// Object model.  An fmgpu_index is the reference's one immutable, @ThreadSafe FmIndex (fm/FmIndex.java:82): the device
// layout is uploaded once and REPLICATED on every device the loader was given (upload to the first, cudaMemcpyPeer over
// NVLink to the others).  Each replica owns a small pool of call contexts (streams, events, scratch buffers); a batch call
// leases one context per replica it uses, so concurrent callers of one handle run concurrently, and calls that share a
// context are ordered on the device by an event (the *_device entry points return before their work has finished).
// Host-pointer batch calls cut the caller's batch into one contiguous slice per replica and write disjoint ranges of the
// caller's outputs.
#include <cuda_runtime.h>

#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fmgpu.h"
#include "flatten.hpp"
#include "host_pack.hpp"
#include "jstream.hpp"
#include "kernels.cuh"
#include "kernels_lf.cuh"
#include "kernels_locate.cuh"
#include "kernels_records.cuh"
#include "kernels_shard.cuh"
#include "kernels_utf8.cuh"
#include "kernels_wavelet.cuh"
#include "kernels_build.cuh"
#include "kernels_dense.cuh"
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

// restores the caller's current device when the entry point returns
struct DeviceRestore {
    int prev = -1;
    DeviceRestore() {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    }
    ~DeviceRestore() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);  // synchronizes with whatever still reads the old buffer
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

// page-locked host staging (packed transport of the host-pointer count call)
struct Pinned {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = n + n / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

enum { CTRL_QUEUE = 0, CTRL_QUEUE2 = 1, CTRL_STATS = 2 /* u64 x FMGPU_N_STATS at word 2.. */, CTRL_WORDS = 64 };

enum { KIND_FM = 0, KIND_WAVELET = 1, KIND_RRR = 2 };

// Everything one in-flight batch call needs on one device.
struct CallCtx {
    static constexpr int PIPE_SLOTS = 8;
    static constexpr int COUNT_CTX = 4;  // compute streams of the chunked host-pointer count call (round-robin, own scratch each)
    bool ready = false;
    cudaStream_t stream = nullptr, copy_stream = nullptr, down_stream = nullptr;
    cudaStream_t cstream[COUNT_CTX - 1] = {nullptr};
    // High-priority streams for the pre-pass kernels of a chunk (descriptors + length sort): they must not queue behind the
    // backward-search CTAs of earlier chunks, or the next search launch is late and the SMs drain.
    cudaStream_t pstream[COUNT_CTX] = {nullptr};
    cudaEvent_t pipe_in[PIPE_SLOTS] = {nullptr}, pipe_out[PIPE_SLOTS] = {nullptr}, pre_done[PIPE_SLOTS] = {nullptr};
    cudaEvent_t done = nullptr;  // end of the most recent call that used this context
    bool done_pending = false;   // ... which may still be running (a *_device call)
    Scratch codes, pats, ctrl, ranges, in_a, in_b, in_c, out_a, out_b, out_c, tmp_a, tmp_b, order, bins, u8conv;
    Pinned stage;
    struct CountCtx {
        Scratch pats, ctrl, order, bins;
    } cctx[COUNT_CTX - 1];
    uint64_t* h_total = nullptr;  // pinned word: total hits of a locate sizing pass
    uint64_t last_launches = 0;
    bool stats_valid = false;
    uint32_t stats_ctx_mask = 1;  // compute contexts whose counters belong to the most recent call

    int create() {
        if (ready) return 0;
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&down_stream, cudaStreamNonBlocking));
        for (int i = 0; i < COUNT_CTX - 1; ++i) CU(cudaStreamCreateWithFlags(&cstream[i], cudaStreamNonBlocking));
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lower = higher priority
        for (int i = 0; i < COUNT_CTX; ++i) CU(cudaStreamCreateWithPriority(&pstream[i], cudaStreamNonBlocking, prio_hi));
        for (int i = 0; i < PIPE_SLOTS; ++i) {
            CU(cudaEventCreateWithFlags(&pipe_in[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&pre_done[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&pipe_out[i], cudaEventDisableTiming));
        }
        CU(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        CU(cudaHostAlloc((void**)&h_total, 64, cudaHostAllocPortable));
        ready = true;
        return 0;
    }
    void destroy() {
        for (Scratch* s : {&codes, &pats, &ctrl, &ranges, &in_a, &in_b, &in_c, &out_a, &out_b, &out_c, &tmp_a, &tmp_b, &order, &bins, &u8conv}) s->release();
        stage.release();
        for (int i = 0; i < COUNT_CTX - 1; ++i) {
            for (Scratch* sc : {&cctx[i].pats, &cctx[i].ctrl, &cctx[i].order, &cctx[i].bins}) sc->release();
            if (cstream[i]) cudaStreamDestroy(cstream[i]);
        }
        for (int i = 0; i < COUNT_CTX; ++i)
            if (pstream[i]) cudaStreamDestroy(pstream[i]);
        for (cudaStream_t s : {stream, copy_stream, down_stream})
            if (s) cudaStreamDestroy(s);
        for (int i = 0; i < PIPE_SLOTS; ++i)
            for (cudaEvent_t e : {pipe_in[i], pipe_out[i], pre_done[i]})
                if (e) cudaEventDestroy(e);
        if (done) cudaEventDestroy(done);
        if (h_total) cudaFreeHost(h_total);
        ready = false;
    }
};

// One host thread per replica of a multi-device handle: the slices of a host-pointer batch call are enqueued (dozens of CUDA
// API calls each) and waited for in parallel instead of one device after the other.
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, quit = false;
    void start() {
        th = std::thread([this] {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                cv.wait(lk, [this] { return has_job || quit; });
                if (quit) return;
                std::function<void()> j = std::move(job);
                lk.unlock();
                j();
                lk.lock();
                has_job = false;
                cv.notify_all();
            }
        });
    }
    void submit(std::function<void()> j) {  // waits for the slot (one job at a time per replica)
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return !has_job; });
        job = std::move(j);
        has_job = true;
        cv.notify_all();
    }
    void wait_idle() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return !has_job; });
    }
    void stop() {
        if (!th.joinable()) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
            cv.notify_all();
        }
        th.join();
    }
};

// One device's copy of the index + its call contexts.
struct Replica {
    int device = 0;
    DevIndex dev{};
    struct Arr {  // a device array of the layout: which DevIndex pointer it backs, its size (for cloning to another device)
        size_t field_offset;
        size_t bytes;
        int layout_slot;
    };
    std::vector<Arr> arrays;
    std::vector<void*> allocs;
    uint64_t layout_bytes[8] = {0};
    uint64_t total_bytes = 0;
    int sm_count = 0;
    int count_ctas = 0, locate_ctas = 0, locate_dense_ctas = 0, extract_ctas = 0, eub_ctas = 0;
    uint64_t dense_bytes = 0;  // dense marks + dsa (also counted in total_bytes)
    // L2 residency of the (block, symbol) cell table: a persisting access-policy window (replica_setup / l2_window)
    size_t l2_window_bytes = 0;
    float l2_hit_ratio = 0.f;
    std::vector<cudaStream_t> l2_streams;  // streams that already carry the window
    size_t tables_smem = 0;
    static constexpr int NCTX = 4;  // concurrent batch calls per device
    std::mutex mu;
    std::condition_variable cv;
    CallCtx ctx[NCTX];
    bool busy[NCTX] = {false, false, false, false};
    int last_ctx = 0;  // context of the most recent call (work counters)
    // optional per-launch timing of the dominant kernels (bench.py's rooflines): one ring of event pairs per kernel kind
    // (FMGPU_KERNEL_COUNT / _LOCATE / _EXTRACT)
    static constexpr int TIMING_SLOTS = 64, TIMING_KINDS = 3;
    cudaEvent_t ev0[TIMING_KINDS][TIMING_SLOTS] = {{nullptr}}, ev1[TIMING_KINDS][TIMING_SLOTS] = {{nullptr}};
    uint64_t timed_calls[TIMING_KINDS] = {0, 0, 0};
};

}  // namespace

struct fmgpu_index {
    int kind = KIND_FM;  // what the handle was loaded from: an FmIndex stream, a bare WaveletFixedBlockBoosting, a bare RrrVector
    int32_t alphabet_length = 0;
    std::vector<std::unique_ptr<Replica>> reps;
    std::vector<std::unique_ptr<Worker>> workers;  // one per replica when there are several
    std::mutex multi_mu;                            // one multi-device call at a time drives the workers
    bool count_stats = false;  // fmgpu_set_stats: kernels with work counters
    bool use_kmer = true;      // fmgpu_set_start_table: patterns start from the q-gram start table when the index has one
    bool use_dense = true;     // fmgpu_set_locate_dense: locate walks end at the device-side dense samples when the index has them
    bool timing = false;
    Replica* primary() const { return reps[0].get(); }
};

namespace {

// A call's claim on one context of one replica.  The constructor makes the replica's device current.
struct Lease {
    Replica* r = nullptr;
    CallCtx* c = nullptr;
    int slot = -1;
    int rc = 0;
    explicit Lease(Replica* rep) : r(rep) {
        {
            std::unique_lock<std::mutex> lk(r->mu);
            for (;;) {
                for (int i = 0; i < Replica::NCTX && slot < 0; ++i)
                    if (!r->busy[i]) slot = i;
                if (slot >= 0) break;
                r->cv.wait(lk);
            }
            r->busy[slot] = true;
            r->last_ctx = slot;
        }
        c = &r->ctx[slot];
        if (cudaSetDevice(r->device) != cudaSuccess) rc = fail(FMGPU_ERR_CUDA, "cannot select device %d", r->device);
        if (!rc) rc = c->create();
    }
    // Orders the call's first stream after the previous call that used this context (it may still be running: the *_device
    // entry points return without synchronizing).
    int begin(cudaStream_t st) {
        if (c->done_pending) CU(cudaStreamWaitEvent(st, c->done, 0));
        return 0;
    }
    // `st` = the stream the call's last piece of work was enqueued on; `synced`: the call has waited for it
    int end(cudaStream_t st, bool synced) {
        if (synced) {
            c->done_pending = false;
            return 0;
        }
        CU(cudaEventRecord(c->done, st));
        c->done_pending = true;
        return 0;
    }
    ~Lease() {
        if (slot >= 0) {
            std::lock_guard<std::mutex> lk(r->mu);
            r->busy[slot] = false;
            r->cv.notify_one();
        }
    }
    Lease(const Lease&) = delete;
    Lease& operator=(const Lease&) = delete;
};

int timing_slot(fmgpu_index* ix, Replica* rp, int kind) {
    if (!ix->timing || !rp->ev0[kind][0]) return -1;
    std::lock_guard<std::mutex> lk(rp->mu);
    return (int)(rp->timed_calls[kind]++ % Replica::TIMING_SLOTS);
}

// slice r of R over n items
inline uint32_t slice_lo(uint32_t n, size_t r, size_t R) { return (uint32_t)((uint64_t)n * r / R); }

// fn(r) for r in [0, R): inline when R == 1, else on the replicas' worker threads, all at once.  Returns the first failure;
// its message becomes the caller's fmgpu_last_error().
template <typename F>
int run_on_replicas(fmgpu_index* ix, size_t R, F&& fn) {
    if (R <= 1 || ix->workers.size() < R) {
        int rc = 0;
        for (size_t r = 0; r < R && !rc; ++r) rc = fn(r);
        return rc;
    }
    std::vector<int> rcs(R, 0);
    std::vector<std::string> errs(R);
    for (size_t r = 0; r < R; ++r)
        ix->workers[r]->submit([&, r] {
            rcs[r] = fn(r);
            if (rcs[r]) errs[r] = g_err;  // the worker thread's own thread-local message
        });
    for (size_t r = 0; r < R; ++r) ix->workers[r]->wait_idle();
    for (size_t r = 0; r < R; ++r)
        if (rcs[r]) {
            g_err = errs[r];
            return rcs[r];
        }
    return 0;
}

template <typename T>
int upload(Replica* rp, const std::vector<T>& v, const T** dptr, int layout_slot) {
    void* p = nullptr;
    const size_t n = v.size() * sizeof(T);
    CU(cudaMalloc(&p, n ? n : 32));
    rp->allocs.push_back(p);
    if (n) CU(cudaMemcpy(p, v.data(), n, cudaMemcpyHostToDevice));
    *dptr = reinterpret_cast<const T*>(p);
    rp->total_bytes += n;
    if (layout_slot >= 0) rp->layout_bytes[layout_slot] += n;
    rp->arrays.push_back({(size_t)((const char*)dptr - (const char*)&rp->dev), n, layout_slot});
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

// the next slot of a kernel kind's timing ring, or -1 when timing is off
int timing_slot(fmgpu_index* ix, Replica* rp, int kind);

// UTF-8 byte patterns: the pre-pass decodes d_bytes[pat_off[i], pat_off[i+1]) into d_chars at the same offsets (kernels_utf8.cuh)
struct Utf8Src {
    const uint8_t* d_bytes;
    uint16_t* d_chars;
    int32_t* d_conv_status;  // per pattern of this call: status of the conversion (0, 9, 10) ...
    int32_t* d_conv_value;   // ... and the offending code point
};

// The backward search reads one cell per rank query from a table that is small (8 bytes per (block, symbol) pair) but shares the
// L2 with the occurrence records streaming through it; ncu showed the cell loads missing L2 about half the time.  The table is
// therefore given a PERSISTING access-policy window on every stream that launches k_count (set-aside sized at load,
// replica_setup): measured -10 % kernel time on the configs[1] batch.  FMGPU_L2_PERSIST=0 in the environment disables it.
void l2_window(Replica* rp, cudaStream_t st) {
    if (!rp->l2_window_bytes) return;
    {
        std::lock_guard<std::mutex> lk(rp->mu);
        for (cudaStream_t s : rp->l2_streams)
            if (s == st) return;
        if (rp->l2_streams.size() >= 64) rp->l2_streams.clear();
        rp->l2_streams.push_back(st);
    }
    cudaStreamAttrValue av{};
    av.accessPolicyWindow.base_ptr = (void*)rp->dev.cells;
    av.accessPolicyWindow.num_bytes = rp->l2_window_bytes;
    av.accessPolicyWindow.hitRatio = rp->l2_hit_ratio;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) (void)cudaGetLastError();
}

// Backward search over n_pat patterns on stream `st`.  `first_of_call` resets the work counters; later
// chunks of the same call only re-arm the work queue.
int count_on_stream(fmgpu_index* ix, Replica* rp, CallCtx* cx, const uint16_t* d_chars, const uint64_t* d_pat_off, uint32_t n_pat,
                    int32_t* d_counts, int32_t* d_status, uint32_t* d_ranges, cudaStream_t st, bool first_of_call = true, int ctx = 0,
                    const Utf8Src* u8 = nullptr, cudaStream_t pre = nullptr, cudaEvent_t pre_ev = nullptr, int threads = CTA_THREADS) {
    // `pre` (optional): stream for the pre-pass kernels, joined into `st` through pre_ev before the search kernel
    if (ix->kind != KIND_FM) return fail(FMGPU_ERR_UNSUPPORTED, "the handle holds no FmIndex (loaded from a bare wavelet / RRR stream)");
    if (u8) d_chars = u8->d_chars;
    Scratch& s_pats = ctx ? cx->cctx[ctx - 1].pats : cx->pats;
    Scratch& s_ctrl = ctx ? cx->cctx[ctx - 1].ctrl : cx->ctrl;
    Scratch& s_order = ctx ? cx->cctx[ctx - 1].order : cx->order;
    Scratch& s_bins = ctx ? cx->cctx[ctx - 1].bins : cx->bins;
    CU(s_pats.reserve(((size_t)n_pat + 2) * sizeof(PatDesc)));
    CU(s_ctrl.reserve(CTRL_WORDS * 4));
    CU(s_order.reserve((size_t)n_pat * 4 + 64));
    CU(s_bins.reserve(LEN_BINS * 4));
    if (!pre) pre = st;
    if (first_of_call) {
        CU(cudaMemsetAsync(s_ctrl.p, 0, CTRL_WORDS * 4, pre));
        if (ctx == 0) {
            cx->last_launches = 0;
            cx->stats_valid = true;
            cx->stats_ctx_mask = 1;
        } else {
            cx->stats_ctx_mask |= 1u << ctx;
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
    const DevIndex& D = rp->dev;
    const int pre_grid = prepass_grid(n_pat, rp->sm_count);
    const uint32_t kq = ix->use_kmer ? D.kmer_q : 0u;  // 0: every pattern starts from its last char
    // descriptors + length histogram, then a counting sort by length so that a warp's 32 patterns run in lockstep
    if (u8)
        k_prepass_utf8<<<pre_grid, 256, 0, pre>>>(u8->d_bytes, d_pat_off, n_pat, D.char2code, u8->d_chars, (PatDesc*)s_pats.p,
                                                 (uint32_t*)s_bins.p, u8->d_conv_status, u8->d_conv_value, kq, D.kmer_stride, D.sigma);
    else
        k_prepass<<<pre_grid, 256, 0, pre>>>(d_chars, d_pat_off, n_pat, D.char2code, (PatDesc*)s_pats.p, (uint32_t*)s_bins.p, kq,
                                             D.kmer_stride, D.sigma);
    k_len_scan<<<1, SCAN_THREADS, 0, pre>>>((uint32_t*)s_bins.p);
    const int sc_grid = prepass_grid(((uint64_t)n_pat + SCATTER_PER_THREAD - 1) / SCATTER_PER_THREAD, rp->sm_count);
    k_len_scatter<<<sc_grid, 256, 0, pre>>>((const PatDesc*)s_pats.p, n_pat, (uint32_t*)s_bins.p, (uint32_t*)s_order.p);
    if (pre != st) {
        CU(cudaEventRecord(pre_ev, pre));
        CU(cudaStreamWaitEvent(st, pre_ev, 0));
    }
    l2_window(rp, st);
    const int slot = timing_slot(ix, rp, FMGPU_KERNEL_COUNT);
    if (slot >= 0) CU(cudaEventRecord(rp->ev0[FMGPU_KERNEL_COUNT][slot], st));
    int grid = rp->count_ctas;
    const int need = (int)(((uint64_t)n_pat + threads - 1) / threads);
    if (need < grid) grid = need;
    if (ix->count_stats)
        k_count<true><<<grid, threads, rp->tables_smem, st>>>(D, d_chars, (const PatDesc*)s_pats.p, (const uint32_t*)s_order.p, n_pat,
                                                                d_counts, d_status, d_ranges, ctrl + CTRL_QUEUE,
                                                                (unsigned long long*)(ctrl + CTRL_STATS));
    else
        k_count<false><<<grid, threads, rp->tables_smem, st>>>(D, d_chars, (const PatDesc*)s_pats.p, (const uint32_t*)s_order.p, n_pat,
                                                                 d_counts, d_status, d_ranges, ctrl + CTRL_QUEUE,
                                                                 (unsigned long long*)(ctrl + CTRL_STATS));
    if (slot >= 0) CU(cudaEventRecord(rp->ev1[FMGPU_KERNEL_COUNT][slot], st));
    cx->last_launches += 4;
    if (u8) {  // a pattern whose conversion throws never reaches the search in the reference: its status wins
        k_utf8_merge<<<(n_pat + 255) / 256, 256, 0, st>>>(u8->d_conv_status, u8->d_conv_value, n_pat, d_counts, d_status, d_ranges ? 0 : 1);
        cx->last_launches += 1;
    }
    CU(cudaGetLastError());
    return 0;
}

constexpr int START_TABLE_LOG2 = 25;  // up to 32 M entries = 256 MB (q = 4 for a 70-symbol log alphabet: 24 M entries)

// q-gram start table (layout.h): the search kernel itself computes, for every q-gram of alphabet codes, the SA range after its
// q chars.  Built once per load on the primary replica (the other replicas receive a copy).
int build_start_table(fmgpu_index* ix, const std::vector<uint16_t>& code2char, const std::vector<uint16_t>& char2code) {
    if (const char* e = getenv("FMGPU_START_TABLE"))
        if (atoi(e) == 0) return 0;
    Replica* rp = ix->primary();
    const uint64_t S = rp->dev.sigma;
    int log2_max = START_TABLE_LOG2;  // FMGPU_START_TABLE_LOG2: largest table, in log2 entries of 8 bytes
    if (const char* e = getenv("FMGPU_START_TABLE_LOG2")) log2_max = atoi(e) >= 2 && atoi(e) <= 27 ? atoi(e) : log2_max;
    {  // no more than ~4 entries per text position: a small index does not get a table larger than itself
        int lb = 2;
        while ((1ull << lb) < (uint64_t)rp->dev.length * 4 && lb < 27) ++lb;
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
        Lease L(rp);
        if (L.rc) return L.rc;
        uint16_t* d_chars = nullptr;
        uint64_t* d_off = nullptr;
        int32_t *d_counts = nullptr, *d_status = nullptr;
        uint32_t* d_ranges = nullptr;
        cudaStream_t st = L.c->stream;
        CU(cudaMalloc((void**)&d_chars, chars.size() * 2));
        CU(cudaMalloc((void**)&d_off, off.size() * 8));
        CU(cudaMalloc((void**)&d_counts, (size_t)n * 4));
        CU(cudaMalloc((void**)&d_status, (size_t)n * 4));
        CU(cudaMalloc((void**)&d_ranges, (size_t)n * 8));
        CU(cudaMemcpyAsync(d_chars, chars.data(), chars.size() * 2, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
        const bool was = ix->use_kmer;
        ix->use_kmer = false;
        int rc = count_on_stream(ix, rp, L.c, d_chars, d_off, n, d_counts, d_status, d_ranges, st);
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
            for (Scratch* sc : {&L.c->pats, &L.c->order}) sc->release();
        L.c->stats_valid = false;
        if (rc) return rc;
        for (uint32_t i = 0; i < n; ++i)
            if (status[i] == 0) table[idx_of[i]] = U32x2{ranges[2 * (size_t)i], ranges[2 * (size_t)i + 1]};
    }
    int rc = upload(rp, table, &rp->dev.kmer, -1);
    if (rc) return rc;
    rp->dev.kmer_q = q;
    rp->dev.kmer_stride = (uint32_t)S;
    return 0;
}

// a device array made on the device itself joins the replica's layout (cloned to the other devices like the uploaded ones)
template <typename T>
void adopt(Replica* rp, const T* p, size_t bytes, const T** dptr, int layout_slot) {
    rp->allocs.push_back((void*)p);
    *dptr = p;
    rp->total_bytes += bytes;
    if (layout_slot >= 0) rp->layout_bytes[layout_slot] += bytes;
    rp->arrays.push_back({(size_t)((const char*)dptr - (const char*)&rp->dev), bytes, layout_slot});
}

constexpr int LOCATE_SAMPLE_RATE_DEFAULT = 8;

// Device-side denser sampling of the SA rows (layout.h, kernels_dense.cuh).  `want` = fmgpu_opts.locate_sample_rate: > 0 the
// requested rate, 0 = FMGPU_LOCATE_SAMPLE_RATE or the default, < 0 = none.  The effective rate is the largest divisor of the
// index's sampleRate that is <= the request (none if that is the sampleRate itself).  Skipped where a shorter walk could
// differ from the reference's: an index on which LF steps can throw (length % 2^20 == 0, quirk Q4) or run in cycles (more than
// 256 symbols, quirk Q1) — there the reference's full walk reports the exception, so the full walk is what runs.  Also skipped
// (silently: it is an accelerator, not a result) when the device has too little free memory.
int build_dense_samples(fmgpu_index* ix, int want) {
    Replica* rp = ix->primary();
    const DevIndex& D = rp->dev;
    if (want == 0) {
        want = LOCATE_SAMPLE_RATE_DEFAULT;
        if (const char* e = getenv("FMGPU_LOCATE_SAMPLE_RATE")) want = atoi(e);
    }
    if (want <= 0 || D.sample_rate < 2 || D.q4 || D.sigma > 256 || D.n_sa == 0 || D.length < 2) return 0;
    uint32_t rate = 0;
    for (uint32_t d = 1; d <= (uint32_t)want && d < D.sample_rate; ++d)
        if (D.sample_rate % d == 0) rate = d;
    if (rate == 0) return 0;
    const uint32_t n_rec = (D.length + DENSE_ROWS_PER_REC - 1) / DENSE_ROWS_PER_REC;
    const uint32_t n_dense = (D.length - 1u) / rate + 1u;  // text positions 0 .. length - 1 that are multiples of rate
    const size_t mark_bytes = (size_t)n_rec * 32, dsa_bytes = (size_t)n_dense * 4, seed_bytes = ((size_t)dense_seed_count(D) + 1) * 4;
    const uint32_t n_blocks = (n_rec + SCAN_BLOCK - 1) / SCAN_BLOCK;
    const size_t scan_bytes = (size_t)n_rec * 4 + ((size_t)n_rec + 1) * 8 + ((size_t)n_blocks + 2) * 8;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return fail(FMGPU_ERR_CUDA, "cudaMemGetInfo failed");
    if ((mark_bytes + dsa_bytes) * 2 + seed_bytes + scan_bytes > free_b) return 0;
    Lease L(rp);
    if (L.rc) return L.rc;
    cudaStream_t st = L.c->stream;
    Rec32* marks = nullptr;
    uint32_t *dsa = nullptr, *seeds = nullptr;
    unsigned int* flag = nullptr;
    char* scan = nullptr;
    auto drop = [&] {
        for (void* p : {(void*)marks, (void*)dsa, (void*)seeds, (void*)flag, (void*)scan})
            if (p) cudaFree(p);
    };
    auto check = [&](cudaError_t e) {
        if (e == cudaSuccess) return 0;
        drop();
        return fail(FMGPU_ERR_CUDA, "dense samples: %s", cudaGetErrorString(e));
    };
    int rc = 0;
    if ((rc = check(cudaMalloc((void**)&marks, mark_bytes))) || (rc = check(cudaMalloc((void**)&dsa, dsa_bytes))) ||
        (rc = check(cudaMalloc((void**)&seeds, seed_bytes))) || (rc = check(cudaMalloc((void**)&flag, 4))) ||
        (rc = check(cudaMalloc((void**)&scan, scan_bytes))))
        return rc;
    int32_t* ones = (int32_t*)scan;
    uint64_t* before = (uint64_t*)(scan + (((size_t)n_rec * 4 + 7) & ~(size_t)7));
    uint64_t* sums = before + n_rec + 1;
    cudaMemsetAsync(marks, 0, mark_bytes, st);
    cudaMemsetAsync(dsa, 0, dsa_bytes, st);
    cudaMemsetAsync(seeds, 0xff, seed_bytes, st);
    cudaMemsetAsync(flag, 0, 4, st);
    const int grid = rp->sm_count * 2;
    const size_t tsmem = tables_smem_bytes(D);
    const uint32_t n_seeds = dense_seed_count(D);
    k_dense_seeds<<<grid, DENSE_THREADS, LOCATE_TAB_WORDS * 4, st>>>(D, seeds, n_seeds);
    k_dense_walk<0><<<grid, DENSE_THREADS, tsmem, st>>>(D, seeds, n_seeds, rate, (uint32_t*)marks, dsa, n_dense, flag);
    k_dense_popc<<<(n_rec + 255) / 256, 256, 0, st>>>(marks, n_rec, ones);
    k_scan_local<<<n_blocks, SCAN_BLOCK, 0, st>>>(ones, n_rec, before, sums);
    k_scan_sums<<<1, SCAN_BLOCK, 0, st>>>(sums, n_blocks, sums + n_blocks);
    k_scan_apply<<<n_blocks, SCAN_BLOCK, 0, st>>>(before, n_rec, sums, sums + n_blocks);
    k_dense_fill<<<(n_rec + 255) / 256, 256, 0, st>>>(marks, n_rec, before);
    k_dense_walk<1><<<grid, DENSE_THREADS, tsmem, st>>>(D, seeds, n_seeds, rate, (uint32_t*)marks, dsa, n_dense, flag);
    uint64_t marked = 0;
    unsigned int failed = 0;
    if ((rc = check(cudaGetLastError())) || (rc = check(cudaMemcpyAsync(&marked, before + n_rec, 8, cudaMemcpyDeviceToHost, st))) ||
        (rc = check(cudaMemcpyAsync(&failed, flag, 4, cudaMemcpyDeviceToHost, st))) || (rc = check(cudaStreamSynchronize(st))))
        return rc;
    cudaFree(seeds);
    cudaFree(flag);
    cudaFree(scan);
    if (failed || marked != n_dense) {  // not the text-order walk of a consistent index: locate keeps the index's own samples
        cudaFree(marks);
        cudaFree(dsa);
        return 0;
    }
    adopt<Rec32>(rp, marks, mark_bytes, &rp->dev.dmarks, -1);
    adopt<uint32_t>(rp, dsa, dsa_bytes, &rp->dev.dsa, -1);
    rp->dense_bytes = mark_bytes + dsa_bytes;
    rp->dev.dense_rate = rate;
    rp->dev.n_dense = n_dense;
    return 0;
}

// per-device setup after the layout is resident: grid sizes, shared-memory opt-ins
int lf_setup(Replica* rp);

int replica_setup(Replica* rp) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, rp->device) != cudaSuccess) return fail(FMGPU_ERR_CUDA, "cudaGetDeviceProperties failed");
    rp->sm_count = prop.multiProcessorCount;
    rp->tables_smem = count_smem_bytes(rp->dev);
    {  // persisting-L2 set-aside for the cell table (l2_window): the table + 25 %, at most half of what the device allows; a
       // larger table gets the fraction of its accesses that fits
        int want_mb = -1;
        if (const char* e = getenv("FMGPU_L2_PERSIST")) want_mb = atoi(e);
        const size_t cells = (size_t)rp->layout_bytes[0];
        size_t set_aside = want_mb > 0 ? (size_t)want_mb << 20 : cells + cells / 4 + (1u << 20);
        const size_t cap = (size_t)prop.persistingL2CacheMaxSize / (want_mb > 0 ? 1 : 2);
        if (set_aside > cap) set_aside = cap;
        rp->l2_window_bytes = 0;
        // the set-aside is one per device and process: it only grows (a small index loaded after a large one must not take
        // the large one's L2 away)
        static std::mutex l2_mu;
        static size_t l2_set_aside[64] = {0};
        bool reserved = false;
        if (want_mb != 0 && cells > 0 && set_aside > 0 && prop.accessPolicyMaxWindowSize > 0) {
            std::lock_guard<std::mutex> lk(l2_mu);
            size_t& cur = l2_set_aside[rp->device & 63];
            if (set_aside <= cur) {
                reserved = true;
            } else if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside) == cudaSuccess) {
                cur = set_aside;
                reserved = true;
            }
        }
        if (reserved) {
            rp->l2_window_bytes = cells < (size_t)prop.accessPolicyMaxWindowSize ? cells : (size_t)prop.accessPolicyMaxWindowSize;
            const double r = (double)set_aside / 1.25 / (double)rp->l2_window_bytes;
            rp->l2_hit_ratio = r >= 1.0 ? 1.0f : (float)r;
        } else {
            (void)cudaGetLastError();
        }
    }
    // the attribute is per function, not per index: always the largest table set any index can stage
    if (cudaFuncSetAttribute((const void*)k_count<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COUNT_SMEM_MAX_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute((const void*)k_count<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COUNT_SMEM_MAX_BYTES) != cudaSuccess)
        return fail(FMGPU_ERR_CUDA, "k_count: cannot reserve %zu bytes of shared memory", rp->tables_smem);
    int g0 = 0, g1 = 0;
    int rc = grid_for((const void*)k_count<false>, rp->sm_count, rp->tables_smem, &g0);
    if (!rc) rc = grid_for((const void*)k_count<true>, rp->sm_count, rp->tables_smem, &g1);
    if (rc) return rc;
    rp->count_ctas = g0 < g1 ? g0 : g1;
    return lf_setup(rp);
}

// the replica whose device holds the caller's device buffer `p` (the *_device entry points)
Replica* replica_of(fmgpu_index* ix, const void* p) {
    if (ix->reps.size() == 1 || !p) return ix->primary();
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice)
        for (auto& r : ix->reps)
            if (r->device == a.device) return r.get();
    (void)cudaGetLastError();
    return nullptr;
}

}  // namespace

#include "api_lf.inc"
#include "api_wavelet.inc"
#include "api_build.inc"
#include "api_shard.inc"

extern "C" {

const char* fmgpu_last_error(void) { return g_err.c_str(); }
const char* fmgpu_version(void) { return "fmgpu 0.2 (sm_100a)"; }

void fmgpu_index_free(fmgpu_index* ix) {
    if (!ix) return;
    for (auto& w : ix->workers) w->stop();
    DeviceRestore keep;
    for (auto& up : ix->reps) {
        Replica* rp = up.get();
        if (cudaSetDevice(rp->device) != cudaSuccess) continue;
        cudaDeviceSynchronize();
        if (rp->l2_window_bytes) cudaCtxResetPersistingL2Cache();  // the cell table's persisting lines
        for (void* p : rp->allocs) cudaFree(p);
        for (CallCtx& c : rp->ctx) c.destroy();
        for (int k = 0; k < Replica::TIMING_KINDS; ++k)
            for (int i = 0; i < Replica::TIMING_SLOTS; ++i) {
                if (rp->ev0[k][i]) cudaEventDestroy(rp->ev0[k][i]);
                if (rp->ev1[k][i]) cudaEventDestroy(rp->ev1[k][i]);
            }
    }
    delete ix;
}

}  // extern "C"

namespace {

// copies the primary replica's device arrays to `device` (cudaMemcpyPeer: NVLink where the GPUs are peers)
int clone_replica(const Replica* src, int device, Replica* dst) {
    dst->device = device;
    dst->dev = src->dev;
    dst->dense_bytes = src->dense_bytes;
    CU(cudaSetDevice(device));
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
        if (e != cudaSuccess) (void)cudaGetLastError();  // already enabled / not supported: the copy below stages instead
    }
    for (const Replica::Arr& a : src->arrays) {
        void* p = nullptr;
        CU(cudaMalloc(&p, a.bytes ? a.bytes : 32));
        dst->allocs.push_back(p);
        const void* from = *reinterpret_cast<void* const*>((const char*)&src->dev + a.field_offset);
        if (a.bytes) CU(cudaMemcpyPeer(p, device, from, src->device, a.bytes));
        *reinterpret_cast<void**>((char*)&dst->dev + a.field_offset) = p;
        dst->total_bytes += a.bytes;
        if (a.layout_slot >= 0) dst->layout_bytes[a.layout_slot] += a.bytes;
        dst->arrays.push_back(a);
    }
    return replica_setup(dst);
}

// Parses a serialized FmIndex / WaveletFixedBlockBoosting / RrrVector (Java stream layout), re-lays it out and uploads it.
int load_common(const uint8_t* buf, size_t len, const fmgpu_opts* opts, fmgpu_index** out, int kind) {
    if (!buf || !out) return fail(FMGPU_ERR_ARG, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(FMGPU_ERR_CUDA, "no CUDA device available (libfmgpu has no CPU fallback)");
    DeviceRestore keep;
    // devices to replicate on: opts->devices[0 .. n_devices), n_devices < 0 = every visible device, 0 = opts->device alone
    std::vector<int> devices;
    const int n_req = opts ? opts->n_devices : 0;
    if (n_req < 0) {
        for (int d = 0; d < ndev; ++d) devices.push_back(d);
    } else if (n_req > 0) {
        if (!opts->devices) return fail(FMGPU_ERR_ARG, "n_devices > 0 without a device list");
        for (int i = 0; i < n_req; ++i) devices.push_back(opts->devices[i]);
    } else {
        int dev = opts ? opts->device : -1;
        if (dev < 0) CU(cudaGetDevice(&dev));
        devices.push_back(dev);
    }
    for (size_t i = 0; i < devices.size(); ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) return fail(FMGPU_ERR_ARG, "device ordinal %d out of range", devices[i]);
        for (size_t j = 0; j < i; ++j)
            if (devices[j] == devices[i]) return fail(FMGPU_ERR_ARG, "device %d listed twice", devices[i]);
    }
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

    if (cudaSetDevice(devices[0]) != cudaSuccess) return fail(FMGPU_ERR_CUDA, "cannot select device %d", devices[0]);
    fmgpu_index* ix = new fmgpu_index();
    ix->kind = kind;
    ix->alphabet_length = F.alphabet_length;
    ix->reps.emplace_back(new Replica());
    Replica* rp = ix->primary();
    rp->device = devices[0];
    rp->dev = F.meta;
    int rc = 0;
    auto up = [&](int r) {
        if (!rc) rc = r;
    };
    up(upload(rp, F.C, &rp->dev.C, -1));
    up(upload(rp, F.char2code, &rp->dev.char2code, -1));
    up(upload(rp, F.code2char, &rp->dev.code2char, -1));
    up(upload(rp, F.sb, &rp->dev.sb, -1));
    up(upload(rp, F.cells, &rp->dev.cells, 0));
    up(upload(rp, F.sectors, &rp->dev.sectors, 1));
    up(upload(rp, F.nodes, &rp->dev.nodes, 2));
    up(upload(rp, F.blocks, &rp->dev.blocks, 3));
    up(upload(rp, F.occ, &rp->dev.occ, 4));
    up(upload(rp, F.sgroups, &rp->dev.sgroups, 5));
    up(upload(rp, F.soffsets, &rp->dev.soffsets, 5));
    {
        const fmgpu_host::RrrTables& RT = fmgpu_host::rrr_tables();
        std::vector<uint16_t> inv(RT.inverse, RT.inverse + 32768), cb(RT.class_base, RT.class_base + 16);
        up(upload(rp, inv, &rp->dev.rrr_inv, -1));
        up(upload(rp, cb, &rp->dev.rrr_cbase, -1));
    }
    up(upload(rp, F.sa, &rp->dev.sa, 6));
    up(upload(rp, F.isa, &rp->dev.isa, 7));
    if (!rc) rc = replica_setup(rp);
    if (!rc) {
        // optional experiment knob: L2 fetch granularity for the random 32-byte record gathers
        const char* g = getenv("FMGPU_L2_FETCH");
        if (g && atoi(g) > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    }
    if (!rc && kind == KIND_FM) rc = build_start_table(ix, F.code2char, F.char2code);
    if (!rc && kind == KIND_FM) rc = build_dense_samples(ix, opts ? opts->locate_sample_rate : 0);
    for (size_t i = 1; !rc && i < devices.size(); ++i) {
        ix->reps.emplace_back(new Replica());
        rc = clone_replica(rp, devices[i], ix->reps.back().get());
    }
    if (rc) {
        fmgpu_index_free(ix);
        return rc;
    }
    if (ix->reps.size() > 1)
        for (size_t i = 0; i < ix->reps.size(); ++i) {
            ix->workers.emplace_back(new Worker());
            ix->workers.back()->start();
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

int32_t fmgpu_input_length(const fmgpu_index* ix) { return ix ? (int32_t)ix->primary()->dev.length : -1; }
int32_t fmgpu_alphabet_length(const fmgpu_index* ix) { return ix ? ix->alphabet_length : -1; }
int32_t fmgpu_sample_rate(const fmgpu_index* ix) { return ix ? (int32_t)ix->primary()->dev.sample_rate : -1; }
int32_t fmgpu_extract_enabled(const fmgpu_index* ix) { return ix ? (int32_t)ix->primary()->dev.extract_enabled : -1; }
int32_t fmgpu_device(const fmgpu_index* ix) { return ix ? ix->primary()->device : -1; }
int32_t fmgpu_num_devices(const fmgpu_index* ix) { return ix ? (int32_t)ix->reps.size() : -1; }
int32_t fmgpu_device_at(const fmgpu_index* ix, int32_t i) { return ix && i >= 0 && (size_t)i < ix->reps.size() ? ix->reps[(size_t)i]->device : -1; }
uint64_t fmgpu_device_bytes(const fmgpu_index* ix) { return ix ? ix->primary()->total_bytes : 0; }
void fmgpu_layout_bytes(const fmgpu_index* ix, uint64_t out8[8]) {
    for (int i = 0; i < 8; ++i) out8[i] = ix ? ix->primary()->layout_bytes[i] : 0;
}

// Page-locked host memory for the host-pointer batch calls: cudaMemcpyAsync only overlaps with the kernels (and reaches the
// PCIe rate) from page-locked buffers.
int32_t fmgpu_host_pack_threads(void) { return (int32_t)fmgpu_host::PackPool::get().threads(); }

int fmgpu_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return fail(FMGPU_ERR_ARG, "null argument");
    CU(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return 0;
}
int fmgpu_host_unregister(void* p) {
    if (!p) return fail(FMGPU_ERR_ARG, "null argument");
    CU(cudaHostUnregister(p));
    return 0;
}
int fmgpu_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(FMGPU_ERR_ARG, "null argument");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return 0;
}
int fmgpu_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return 0;
}

int fmgpu_count_batch_device(fmgpu_index* ix, const uint16_t* d_chars, const uint64_t* d_pat_off, uint64_t total_chars, uint32_t n_pat,
                             int32_t* d_counts_out, int32_t* d_status_out, void* cuda_stream) {
    if (!ix || !d_pat_off || !d_counts_out || (!d_chars && total_chars)) return fail(FMGPU_ERR_ARG, "null argument");
    Replica* rp = replica_of(ix, d_pat_off);
    if (!rp) return fail(FMGPU_ERR_ARG, "the device buffers are not on a device that holds this index");
    DeviceRestore keep;
    Lease L(rp);
    if (L.rc) return L.rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = L.begin(st);
    if (!rc) rc = count_on_stream(ix, rp, L.c, d_chars, d_pat_off, n_pat, d_counts_out, d_status_out, nullptr, st);
    if (!rc) rc = L.end(st, false);
    return rc;
}

}  // extern "C"

namespace {

// Host-pointer count call on ONE replica, patterns [lo, hi) of the caller's batch — enqueue phase.  The batch slice is cut
// into chunks: the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the kernels of chunk k (copy stream, COUNT_CTX
// compute streams round-robin, download stream, events in between).  A launch lasts at least as long as its longest
// pattern's dependent chain (~0.2 ms), so consecutive chunks run on different compute streams and the next chunk's CTAs fill
// the SMs while the previous chunk's last warps drain.  Chunk chars land at their offsets relative to the slice's first char.
int count_host_enqueue(fmgpu_index* ix, Lease& L, const void* in, size_t unit, const uint64_t* pat_off, uint32_t lo_pat, uint32_t hi_pat,
                       int32_t* counts_out, int32_t* status_out, bool allow_pack) {
    Replica* rp = L.r;
    CallCtx* cx = L.c;
    CU(cudaSetDevice(rp->device));
    const bool utf8 = unit == 1;
    const uint32_t n_pat = hi_pat - lo_pat;
    const uint64_t base = pat_off[lo_pat];
    const uint64_t total = pat_off[hi_pat] - base;
    cudaStream_t st = cx->stream, cp = cx->copy_stream;
    CU(cx->in_a.reserve((size_t)total * 2 + 64));
    CU(cx->in_b.reserve(((size_t)n_pat + 1) * 8));
    CU(cx->out_a.reserve((size_t)n_pat * 4 + 64));
    CU(cx->out_b.reserve((size_t)n_pat * 4 + 64));
    if (utf8) {
        CU(cx->codes.reserve((size_t)total + 64));
        CU(cx->u8conv.reserve((size_t)n_pat * 8 + 64));
    }
    int rc = L.begin(cp);
    if (rc) return rc;
    uint32_t min_chunk = 125000;
    if (const char* e = getenv("FMGPU_PIPE_CHUNK")) min_chunk = (uint32_t)atoi(e) > 0 ? (uint32_t)atoi(e) : min_chunk;
    int n_ctx = CallCtx::COUNT_CTX;
    if (const char* e = getenv("FMGPU_PIPE_STREAMS")) n_ctx = atoi(e) >= 1 && atoi(e) <= CallCtx::COUNT_CTX ? atoi(e) : n_ctx;
    uint32_t n_chunks = n_pat / min_chunk;
    if (n_chunks > (uint32_t)CallCtx::PIPE_SLOTS) n_chunks = CallCtx::PIPE_SLOTS;
    if (n_chunks < 1) n_chunks = 1;
    // search CTAs of the chunked call are a little smaller than CTA_THREADS so that two of them leave registers and thread
    // slots on the SM for the (small) pre-pass CTAs of the next chunks
    int pipe_threads = n_chunks > 1 ? PIPE_CTA_THREADS : CTA_THREADS;
    if (const char* e = getenv("FMGPU_PIPE_THREADS")) pipe_threads = atoi(e) >= 32 && atoi(e) <= CTA_THREADS ? (atoi(e) / 32) * 32 : pipe_threads;
    // The device copies of the chars / offsets are indexed with the caller's ABSOLUTE offsets (no rebasing of pat_off): the
    // buffers hold the slice only, so their base pointers are shifted back by the slice's first offset.  Nothing below the
    // slice is ever dereferenced.
    uint16_t* d_chars = (uint16_t*)cx->in_a.p - base;
    uint8_t* d_in = (utf8 ? (uint8_t*)cx->codes.p : (uint8_t*)cx->in_a.p) - base * unit;
    uint64_t* d_off = (uint64_t*)cx->in_b.p - lo_pat;
    int32_t* d_counts = (int32_t*)cx->out_a.p - lo_pat;
    int32_t* d_status = (int32_t*)cx->out_b.p - lo_pat;
    int32_t* d_conv = (int32_t*)cx->u8conv.p;
    const uint8_t* h_in = (const uint8_t*)in;
    // FMGPU_PIPE_TRACE=1: per-chunk timeline of the call on stderr (timing events; diagnostic only)
    const bool trace = getenv("FMGPU_PIPE_TRACE") != nullptr;
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    double h_up[CallCtx::PIPE_SLOTS] = {0}, h_kern[CallCtx::PIPE_SLOTS] = {0}, h_submit = 0;
    cudaEvent_t t0 = nullptr, t_in0[CallCtx::PIPE_SLOTS], t_in[CallCtx::PIPE_SLOTS], t_k0[CallCtx::PIPE_SLOTS], t_k1[CallCtx::PIPE_SLOTS], t_out[CallCtx::PIPE_SLOTS];
    if (trace) {
        CU(cudaEventCreate(&t0));
        for (uint32_t k = 0; k < n_chunks; ++k) {
            CU(cudaEventCreate(&t_in0[k]));
            CU(cudaEventCreate(&t_in[k]));
            CU(cudaEventCreate(&t_k0[k]));
            CU(cudaEventCreate(&t_k1[k]));
            CU(cudaEventCreate(&t_out[k]));
        }
        CU(cudaEventRecord(t0, cp));
    }
    // Packed transport (host_pack.hpp): char[] chunks whose chars all fit a byte cross PCIe as bytes + chunk-relative uint32
    // offsets, narrowed by the pack pool into this context's page-locked staging buffer while earlier chunks are on the wire.
    // FMGPU_HOST_PACK=0 sends the caller's arrays as they are (they should then be page-locked).
    // FMGPU_HOST_PACK_MIN: smallest slice (chars) that is packed (tests force the path on small batches).
    bool pack_enabled = true;
    const bool pack_help = !(getenv("FMGPU_PACK_HELP") && atoi(getenv("FMGPU_PACK_HELP")) == 0);  // the caller's thread packs too
    uint64_t pack_min = 1u << 20;
    if (const char* e = getenv("FMGPU_HOST_PACK")) pack_enabled = atoi(e) != 0;
    if (const char* e = getenv("FMGPU_HOST_PACK_MIN")) pack_min = (uint64_t)atoll(e);
    const bool pack = allow_pack && pack_enabled && !utf8 && total >= pack_min && total > 0 && total < (1ull << 32) && fmgpu_host::PackPool::get().threads() > 0;
    std::vector<std::atomic<uint32_t>> wide(pack ? n_chunks : 0);  // per chunk: OR of its chars (> 0xFF: the chunk is sent as it is)
    struct JobGuard {  // whatever way this function is left, the pool must be done with the buffers above first
        std::shared_ptr<fmgpu_host::PackJob> j;
        ~JobGuard() {
            if (j) j->wait_all();
        }
    } guard;
    std::shared_ptr<fmgpu_host::PackJob>& job = guard.j;
    uint8_t* h_bytes = nullptr;
    uint32_t* h_off32 = nullptr;
    uint8_t* d_bytes = nullptr;
    uint32_t* d_off32 = nullptr;
    constexpr uint32_t PACK_PARTS = 16;
    if (pack) {
        const size_t bytes_cap = ((size_t)total + 63) & ~(size_t)63;
        CU(cx->stage.reserve(bytes_cap + ((size_t)n_pat + n_chunks + 1) * 4));
        CU(cx->codes.reserve(bytes_cap + 64));
        CU(cx->in_c.reserve(((size_t)n_pat + n_chunks + 1) * 4));
        h_bytes = (uint8_t*)cx->stage.p;
        h_off32 = (uint32_t*)(h_bytes + bytes_cap);
        d_bytes = (uint8_t*)cx->codes.p;
        d_off32 = (uint32_t*)cx->in_c.p;
        for (auto& w : wide) w.store(0, std::memory_order_relaxed);
        const uint16_t* h16 = (const uint16_t*)in;
        std::atomic<uint32_t>* wide_p = wide.data();
        job = fmgpu_host::PackPool::get().submit(n_chunks, PACK_PARTS, [=](uint32_t k, uint32_t part) {
            const uint32_t lo = lo_pat + (uint32_t)((uint64_t)n_pat * k / n_chunks), hi = lo_pat + (uint32_t)((uint64_t)n_pat * (k + 1) / n_chunks);
            const uint64_t c0 = pat_off[lo], c1 = pat_off[hi];
            const uint64_t a = c0 + (c1 - c0) * part / PACK_PARTS, b = c0 + (c1 - c0) * (part + 1) / PACK_PARTS;
            const uint32_t m = fmgpu_host::narrow_u16(h16 + a, h_bytes + (a - base), (size_t)(b - a));
            if (m > 0xffu) wide_p[k].fetch_or(m, std::memory_order_relaxed);
            // the chunk's offsets, relative to its first char: hi - lo + 1 entries at slot (lo - lo_pat) + k
            const uint32_t n_off = hi - lo + 1;
            const uint32_t i0 = (uint32_t)((uint64_t)n_off * part / PACK_PARTS), i1 = (uint32_t)((uint64_t)n_off * (part + 1) / PACK_PARTS);
            uint32_t* dst = h_off32 + (lo - lo_pat) + k;
            for (uint32_t i = i0; i < i1; ++i) _mm_stream_si32((int*)(dst + i), (int)(uint32_t)(pat_off[lo + i] - c0));  // (non-temporal, as the bytes)
            _mm_sfence();
        });
    }
    // Uploads run ahead of the kernels: before the kernels of chunk k are enqueued, every chunk that is ready (all of them for
    // the direct path; with the packed transport those the pool has finished, and chunk k itself once it has) is put on the
    // copy stream.
    bool narrow_k[CallCtx::PIPE_SLOTS] = {false};
    auto upload_chunk = [&](uint32_t k) -> int {
        const uint32_t lo = lo_pat + (uint32_t)((uint64_t)n_pat * k / n_chunks), hi = lo_pat + (uint32_t)((uint64_t)n_pat * (k + 1) / n_chunks);
        const uint64_t c0 = pat_off[lo], c1 = pat_off[hi];
        if (pack) {
            if (pack_help) job->wait_group(k);
            else
                while (job->done[k].load(std::memory_order_acquire) < PACK_PARTS) std::this_thread::yield();
            narrow_k[k] = wide[k].load(std::memory_order_relaxed) <= 0xffu;
        }
        if (trace) CU(cudaEventRecord(t_in0[k], cp));
        if (narrow_k[k]) {
            const uint32_t slot = (lo - lo_pat) + k, n_off = hi - lo + 1;
            if (c1 > c0) CU(cudaMemcpyAsync(d_bytes + (c0 - base), h_bytes + (c0 - base), (size_t)(c1 - c0), cudaMemcpyHostToDevice, cp));
            CU(cudaMemcpyAsync(d_off32 + slot, h_off32 + slot, (size_t)n_off * 4, cudaMemcpyHostToDevice, cp));
        } else {
            if (c1 > c0) CU(cudaMemcpyAsync(d_in + c0 * unit, h_in + c0 * unit, (size_t)(c1 - c0) * unit, cudaMemcpyHostToDevice, cp));
            CU(cudaMemcpyAsync(d_off + lo, pat_off + lo, ((size_t)(hi - lo) + 1) * 8, cudaMemcpyHostToDevice, cp));
        }
        CU(cudaEventRecord(cx->pipe_in[k], cp));
        if (trace) CU(cudaEventRecord(t_in[k], cp));
        if (trace) h_up[k] = host_ms();
        return 0;
    };
    auto download_chunk = [&](uint32_t k) -> int {
        const uint32_t lo = lo_pat + (uint32_t)((uint64_t)n_pat * k / n_chunks), hi = lo_pat + (uint32_t)((uint64_t)n_pat * (k + 1) / n_chunks);
        CU(cudaStreamWaitEvent(cx->down_stream, cx->pipe_out[k], 0));
        if (hi > lo) {
            CU(cudaMemcpyAsync(counts_out + lo, d_counts + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, cx->down_stream));
            if (status_out)
                CU(cudaMemcpyAsync(status_out + lo, d_status + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, cx->down_stream));
        }
        if (trace) CU(cudaEventRecord(t_out[k], cx->down_stream));
        return 0;
    };
    if (trace) h_submit = host_ms();
    if (pack && getenv("FMGPU_PACK_SYNC")) job->wait_all();  // diagnostic: pack everything before the first upload
    uint32_t next_up = 0;
    for (uint32_t k = 0; k < n_chunks; ++k) {
        while (next_up < n_chunks && (next_up <= k || !pack || job->done[next_up].load(std::memory_order_acquire) >= PACK_PARTS)) {
            if (int e = upload_chunk(next_up)) return e;
            ++next_up;
        }
        const uint32_t lo = lo_pat + (uint32_t)((uint64_t)n_pat * k / n_chunks), hi = lo_pat + (uint32_t)((uint64_t)n_pat * (k + 1) / n_chunks);
        const int ctx = (int)(k % (uint32_t)n_ctx);
        cudaStream_t cs = ctx ? cx->cstream[ctx - 1] : st;
        cudaStream_t ps = cx->pstream[ctx];
        // the pre-pass of chunk k runs on the context's high-priority stream as soon as the chunk has arrived and the previous
        // search of this context (chunk k - n_ctx, same scratch buffers) is done
        CU(cudaStreamWaitEvent(ps, cx->pipe_in[k], 0));
        if (k >= (uint32_t)n_ctx) CU(cudaStreamWaitEvent(ps, cx->pipe_out[k - n_ctx], 0));
        Utf8Src u8{d_in, d_chars, d_conv + (lo - lo_pat), d_conv + n_pat + (lo - lo_pat)};
        if (trace) CU(cudaEventRecord(t_k0[k], ps));
        if (narrow_k[k]) {  // bytes + relative offsets -> the chars and offsets the pre-pass reads
            const uint64_t c0 = pat_off[lo], c1 = pat_off[hi];
            const uint32_t slot = (lo - lo_pat) + k, n_off = hi - lo + 1;
            const uint64_t work = (c1 - c0) > n_off ? (c1 - c0) : n_off;
            k_unpack_narrow<<<prepass_grid(work / 2 + 1, rp->sm_count), 256, 0, ps>>>(d_bytes + (c0 - base), c1 - c0, d_chars + c0, d_off32 + slot, n_off,
                                                                                   c0, d_off + lo);
        }
        rc = count_on_stream(ix, rp, cx, d_chars, d_off + lo, hi - lo, d_counts + lo, d_status + lo, nullptr, cs, k < (uint32_t)n_ctx, ctx,
                             utf8 ? &u8 : nullptr, ps, cx->pre_done[k], pipe_threads);
        if (rc) return rc;
        if (narrow_k[k]) cx->last_launches += 1;
        CU(cudaEventRecord(cx->pipe_out[k], cs));
        if (trace) CU(cudaEventRecord(t_k1[k], cs));
        if (int e = download_chunk(k)) return e;
        if (trace) h_kern[k] = host_ms();
    }
    if (trace) {
        const double h_enq = host_ms();
        CU(cudaStreamSynchronize(cx->down_stream));
        fprintf(stderr, "[fmgpu trace] host: pack job submitted %.3f ms, all enqueued %.3f ms, device done %.3f ms after entry (packed transport %d)\n",
                h_submit, h_enq, host_ms(), (int)pack);
        for (uint32_t k = 0; k < n_chunks; ++k)
            fprintf(stderr, "[fmgpu trace] host: chunk %u upload enqueued %.3f ms, kernels enqueued %.3f ms\n", k, h_up[k], h_kern[k]);
        for (uint32_t k = 0; k < n_chunks; ++k) {
            float a0 = 0, a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime(&a0, t0, t_in0[k]);
            cudaEventElapsedTime(&a, t0, t_in[k]);
            cudaEventElapsedTime(&b, t0, t_k0[k]);
            cudaEventElapsedTime(&c, t0, t_k1[k]);
            cudaEventElapsedTime(&d, t0, t_out[k]);
            fprintf(stderr, "[fmgpu trace] device %d chunk %u: h2d %.3f .. %.3f ms, kernels %.3f .. %.3f ms, d2h done %.3f ms\n", rp->device, k, a0, a, b,
                    c, d);
            cudaEventDestroy(t_in0[k]);
            cudaEventDestroy(t_in[k]);
            cudaEventDestroy(t_k0[k]);
            cudaEventDestroy(t_k1[k]);
            cudaEventDestroy(t_out[k]);
        }
        cudaEventDestroy(t0);
    }
    return 0;
}

int count_host_wait(Lease& L) {
    CallCtx* cx = L.c;
    CU(cudaSetDevice(L.r->device));
    CU(cudaStreamSynchronize(cx->down_stream));
    CU(cudaStreamSynchronize(cx->copy_stream));
    CU(cudaStreamSynchronize(cx->stream));
    for (int i = 0; i < CallCtx::COUNT_CTX - 1; ++i) CU(cudaStreamSynchronize(cx->cstream[i]));
    for (int i = 0; i < CallCtx::COUNT_CTX; ++i) CU(cudaStreamSynchronize(cx->pstream[i]));
    return L.end(cx->stream, true);
}

// Host-pointer count call for char[] patterns (unit 2) or UTF-8 byte patterns (unit 1): one contiguous slice of the batch per
// replica, all slices in flight at once, disjoint ranges of the caller's outputs.
int count_host(fmgpu_index* ix, const void* in, size_t unit, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts_out, int32_t* status_out) {
    if (!ix || !pat_off || !counts_out) return fail(FMGPU_ERR_ARG, "null argument");
    if (pat_off[n_pat] > pat_off[0] && !in) return fail(FMGPU_ERR_ARG, "null argument");
    DeviceRestore keep;
    const size_t R = n_pat >= 2 * ix->reps.size() ? ix->reps.size() : 1;
    std::unique_lock<std::mutex> multi(ix->multi_mu, std::defer_lock);
    if (R > 1) multi.lock();
    return run_on_replicas(ix, R, [&](size_t r) -> int {
        Lease L(ix->reps[r].get());
        if (L.rc) return L.rc;
        // (a multi-replica call feeds several PCIe links at once: together they outrun the host's packing rate, so it goes direct)
        int rc = count_host_enqueue(ix, L, in, unit, pat_off, slice_lo(n_pat, r, R), slice_lo(n_pat, r + 1, R), counts_out, status_out, R == 1);
        const int w = count_host_wait(L);  // wait for whatever was enqueued, also after a failure
        return rc ? rc : w;
    });
}

}  // namespace

extern "C" {

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
    Replica* rp = replica_of(ix, d_pat_off);
    if (!rp) return fail(FMGPU_ERR_ARG, "the device buffers are not on a device that holds this index");
    DeviceRestore keep;
    Lease L(rp);
    if (L.rc) return L.rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CU(L.c->in_a.reserve((size_t)total_bytes * 2 + 64));
    CU(L.c->u8conv.reserve((size_t)n_pat * 8 + 64));
    Utf8Src u8{d_bytes, (uint16_t*)L.c->in_a.p, (int32_t*)L.c->u8conv.p, (int32_t*)L.c->u8conv.p + n_pat};
    int rc = L.begin(st);
    if (!rc) rc = count_on_stream(ix, rp, L.c, nullptr, d_pat_off, n_pat, d_counts_out, d_status_out, nullptr, st, true, 0, &u8);
    if (!rc) rc = L.end(st, false);
    return rc;
}

int fmgpu_set_start_table(fmgpu_index* ix, int enable) {
    if (!ix) return fail(FMGPU_ERR_ARG, "null argument");
    ix->use_kmer = enable != 0;
    return 0;
}
int32_t fmgpu_start_table_q(const fmgpu_index* ix) { return ix ? (int32_t)ix->primary()->dev.kmer_q : -1; }

int fmgpu_set_locate_dense(fmgpu_index* ix, int enable) {
    if (!ix) return fail(FMGPU_ERR_ARG, "null handle");
    ix->use_dense = enable != 0;
    return 0;
}
int32_t fmgpu_locate_sample_rate(const fmgpu_index* ix) {
    if (!ix) return -1;
    const DevIndex& D = ix->primary()->dev;
    return (int32_t)(D.dense_rate ? D.dense_rate : D.sample_rate);
}
uint64_t fmgpu_dense_sample_bytes(const fmgpu_index* ix) { return ix ? ix->primary()->dense_bytes : 0; }

int fmgpu_set_stats(fmgpu_index* ix, int enable) {
    if (!ix) return fail(FMGPU_ERR_ARG, "null argument");
    ix->count_stats = enable != 0;
    return 0;
}

int fmgpu_set_timing(fmgpu_index* ix, int enable) {
    if (!ix) return fail(FMGPU_ERR_ARG, "null argument");
    DeviceRestore keep;
    for (auto& up : ix->reps) {
        Replica* rp = up.get();
        CU(cudaSetDevice(rp->device));
        std::lock_guard<std::mutex> lk(rp->mu);
        if (enable && !rp->ev0[0][0])
            for (int k = 0; k < Replica::TIMING_KINDS; ++k)
                for (int i = 0; i < Replica::TIMING_SLOTS; ++i) {
                    CU(cudaEventCreate(&rp->ev0[k][i]));
                    CU(cudaEventCreate(&rp->ev1[k][i]));
                }
        for (int k = 0; k < Replica::TIMING_KINDS; ++k) rp->timed_calls[k] = 0;
    }
    ix->timing = enable != 0;
    return 0;
}

int fmgpu_kernel_ms(fmgpu_index* ix, int32_t kind, uint32_t device_index, uint32_t calls_back, float* ms_out) {
    if (!ix || !ms_out || kind < 0 || kind >= Replica::TIMING_KINDS || device_index >= ix->reps.size())
        return fail(FMGPU_ERR_ARG, "null argument / unknown kernel kind / device index out of range");
    DeviceRestore keep;
    Replica* rp = ix->reps[device_index].get();
    CU(cudaSetDevice(rp->device));
    if (!rp->ev0[kind][0] || calls_back >= rp->timed_calls[kind] || calls_back >= (uint32_t)Replica::TIMING_SLOTS)
        return fail(FMGPU_ERR_ARG, "no timing recorded for that launch");
    const int slot = (int)((rp->timed_calls[kind] - 1 - calls_back) % Replica::TIMING_SLOTS);
    CU(cudaEventSynchronize(rp->ev1[kind][slot]));
    CU(cudaEventElapsedTime(ms_out, rp->ev0[kind][slot], rp->ev1[kind][slot]));
    return 0;
}
int fmgpu_search_kernel_ms(fmgpu_index* ix, uint32_t calls_back, float* ms_out) {
    return fmgpu_kernel_ms(ix, FMGPU_KERNEL_COUNT, 0, calls_back, ms_out);
}

int fmgpu_last_stats(fmgpu_index* ix, uint64_t out8[8]) {
    uint64_t v[FMGPU_N_STATS];
    if (!out8) return fail(FMGPU_ERR_ARG, "null argument");
    const int rc = fmgpu_last_stats_ex(ix, v, FMGPU_N_STATS);
    if (!rc) memcpy(out8, v, 8 * sizeof(uint64_t));
    return rc;
}

// counters of the most recent call of every replica (a multi-device call ran one slice on each), summed
int fmgpu_last_stats_ex(fmgpu_index* ix, uint64_t* out8, uint32_t n_out) {
    if (!ix || !out8 || n_out < FMGPU_N_STATS) return fail(FMGPU_ERR_ARG, "null argument or fewer than FMGPU_N_STATS slots");
    DeviceRestore keep;
    memset(out8, 0, FMGPU_N_STATS * sizeof(uint64_t));
    for (auto& up : ix->reps) {
        Replica* rp = up.get();
        CallCtx* cx;
        {
            std::lock_guard<std::mutex> lk(rp->mu);
            cx = &rp->ctx[rp->last_ctx];
        }
        if (!cx->ready || !cx->stats_valid || !cx->ctrl.p) continue;
        CU(cudaSetDevice(rp->device));
        CU(cudaDeviceSynchronize());
        for (int c = 0; c < CallCtx::COUNT_CTX; ++c) {
            if (!(cx->stats_ctx_mask & (1u << c))) continue;
            void* p = c ? cx->cctx[c - 1].ctrl.p : cx->ctrl.p;
            if (!p) continue;
            uint32_t words[CTRL_WORDS];
            CU(cudaMemcpy(words, p, sizeof words, cudaMemcpyDeviceToHost));
            uint64_t v[FMGPU_N_STATS];
            memcpy(v, words + CTRL_STATS, FMGPU_N_STATS * sizeof(uint64_t));
            for (int i = 0; i < FMGPU_N_STATS; ++i)
                if (i != 5) out8[i] += v[i];
        }
        out8[5] += cx->last_launches;
    }
    return 0;
}

}  // extern "C"
