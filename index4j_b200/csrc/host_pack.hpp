// Host-side transport packing for the host-pointer count call (fmgpu.cu, count_host_enqueue).
//
// fmgpu_count_batch is PCIe-bound: 68 of the 76 MB it uploads per 1 M patterns are the UTF-16 chars themselves.  Log text is
// almost always Latin-1, so a pool of host threads narrows every chunk of the caller's char[] to one byte per char (and its
// uint64 offsets to chunk-relative uint32) into the library's own page-locked staging buffer while the previous chunk is on the
// wire; the device widens them again (k_unpack_narrow).  A chunk that holds a char above 0xFF is sent as it is.  Side effect: the
// caller's arrays need not be page-locked for this path — the CPU reads them, the DMA engine reads the staging buffer.
// Measured on the B200 box (16 vCPUs): 13.5 GB/s of char[] per thread, 68 GB/s with 8 threads — above the 55 GB/s of the PCIe link;
// fmgpu_count_batch of 1 M patterns: 0.77 G patterns/s packed against 0.58 G/s direct.
#pragma once
#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace fmgpu_host {

// d[i] = (uint8_t)s[i]; returns the OR of all s[i] (> 0xFF: the range does not fit bytes and d is garbage).
// The bytes are written with NON-TEMPORAL stores: the next reader is the GPU's DMA engine, and lines that sit dirty in the
// cores' caches are slow for it to fetch (measured: the last chunk of a call, which nothing had evicted yet, uploaded at 8 GB/s).
__attribute__((target("avx2"))) inline uint32_t narrow_u16_avx2(const uint16_t* s, uint8_t* d, size_t n) {
    uint32_t t = 0;
    size_t i = 0;
    for (; i < n && ((uintptr_t)(d + i) & 31u); ++i) {  // head: up to the destination's 32-byte boundary
        t |= s[i];
        d[i] = (uint8_t)s[i];
    }
    __m256i acc = _mm256_setzero_si256();
    for (; i + 32 <= n; i += 32) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s + i));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(s + i + 16));
        acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
        _mm256_stream_si256((__m256i*)(d + i), _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xD8));
    }
    for (; i < n; ++i) {
        t |= s[i];
        d[i] = (uint8_t)s[i];
    }
    alignas(32) uint16_t tmp[16];
    _mm256_store_si256((__m256i*)tmp, acc);
    for (int k = 0; k < 16; ++k) t |= tmp[k];
    _mm_sfence();
    return t;
}
inline uint32_t narrow_u16_plain(const uint16_t* s, uint8_t* d, size_t n) {
    uint32_t t = 0;
    for (size_t i = 0; i < n; ++i) {
        t |= s[i];
        d[i] = (uint8_t)s[i];
    }
    return t;
}
inline uint32_t narrow_u16(const uint16_t* s, uint8_t* d, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    return avx2 ? narrow_u16_avx2(s, d, n) : narrow_u16_plain(s, d, n);
}

// A job = n_groups x parts_per_group independent parts, taken in order; the submitter waits per group (a group = one chunk of a
// batch call: it is uploaded as soon as its parts are done, while the pool works on the next group).
struct PackJob {
    std::function<void(uint32_t group, uint32_t part)> fn;
    uint32_t n_groups = 0, parts_per_group = 1;
    std::atomic<uint32_t> next{0};
    std::unique_ptr<std::atomic<uint32_t>[]> done;  // per group
    bool take_one() {  // runs one part; false when none is left
        const uint32_t p = next.fetch_add(1, std::memory_order_relaxed);
        if (p >= n_groups * parts_per_group) return false;
        const uint32_t g = p / parts_per_group;
        fn(g, p % parts_per_group);
        done[g].fetch_add(1, std::memory_order_release);
        return true;
    }
    bool exhausted() const { return next.load(std::memory_order_relaxed) >= n_groups * parts_per_group; }
    // the submitter helps while it waits (so a job also completes without any pool thread)
    void wait_group(uint32_t g) {
        while (done[g].load(std::memory_order_acquire) < parts_per_group)
            if (!take_one()) std::this_thread::yield();
    }
    void wait_all() {
        for (uint32_t g = 0; g < n_groups; ++g) wait_group(g);
    }
};

class PackPool {
public:
    static PackPool& get() {
        static PackPool pool;
        return pool;
    }
    int threads() const { return (int)th_.size(); }
    std::shared_ptr<PackJob> submit(uint32_t n_groups, uint32_t parts_per_group, std::function<void(uint32_t, uint32_t)> fn) {
        auto job = std::make_shared<PackJob>();
        job->fn = std::move(fn);
        job->n_groups = n_groups;
        job->parts_per_group = parts_per_group ? parts_per_group : 1;
        job->done.reset(new std::atomic<uint32_t>[n_groups ? n_groups : 1]);
        for (uint32_t g = 0; g < n_groups; ++g) job->done[g].store(0, std::memory_order_relaxed);
        if (!th_.empty() && n_groups) {
            {
                std::lock_guard<std::mutex> lk(mu_);
                queue_.push_back(job);
            }
            cv_.notify_all();
        }
        return job;
    }
    ~PackPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }

private:
    // Pool size: half the hardware threads, at most 8 — divided by LOCAL_WORLD_SIZE when several processes share the box (one
    // process per GPU under torchrun).  Packing only pays with enough threads to outrun the PCIe link (measured on the B200 box:
    // 4 threads 0.60 G patterns/s = the direct path, 8 threads 0.77 G/s, 12-15 threads no better), so a budget below 8 threads
    // means no pool and no packing.  FMGPU_PACK_THREADS overrides (0 = off).
    PackPool() {
        int share = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) share = atoi(e) > 0 ? atoi(e) : 1;
        int n = (int)std::thread::hardware_concurrency() / 2 / share;
        if (n > 8) n = 8;
        if (n < 8) n = 0;
        if (const char* e = getenv("FMGPU_PACK_THREADS")) n = atoi(e);
        if (n < 0) n = 0;
        if (n > 64) n = 64;
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { worker(); });
    }
    void worker() {
        for (;;) {
            std::shared_ptr<PackJob> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return quit_ || !queue_.empty(); });
                if (quit_) return;
                job = queue_.front();
                if (job->exhausted()) {
                    queue_.pop_front();
                    continue;
                }
            }
            while (job->take_one()) {
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<PackJob>> queue_;
    bool quit_ = false;
};

}  // namespace fmgpu_host
