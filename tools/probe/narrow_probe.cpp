#include <immintrin.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <thread>
#include <chrono>
#include <cstring>
static uint32_t narrow(const uint16_t* s, uint8_t* d, size_t n) {
    __m256i acc = _mm256_setzero_si256();
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(s + i));
        __m256i b = _mm256_loadu_si256((const __m256i*)(s + i + 16));
        acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
        __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xD8);
        _mm256_storeu_si256((__m256i*)(d + i), p);
    }
    uint32_t t = 0;
    for (; i < n; ++i) { t |= s[i]; d[i] = (uint8_t)s[i]; }
    alignas(32) uint16_t tmp[16]; _mm256_store_si256((__m256i*)tmp, acc);
    for (int k = 0; k < 16; ++k) t |= tmp[k];
    return t;
}
int main(int argc, char** argv) {
    int T = argc > 1 ? atoi(argv[1]) : 8;
    size_t n = 34000000;
    std::vector<uint16_t> src(n); for (size_t i = 0; i < n; ++i) src[i] = 32 + (i * 7) % 90;
    std::vector<uint8_t> dst(n);
    for (int rep = 0; rep < 5; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th; std::vector<uint32_t> r(T);
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] { size_t a = n * t / T, b = n * (t + 1) / T; r[t] = narrow(src.data() + a, dst.data() + a, b - a); });
        for (auto& x : th) x.join();
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        printf("T=%d %.3f ms  %.1f GB/s read  flag %x\n", T, ms, n * 2 / ms / 1e6, r[0]);
    }
}
