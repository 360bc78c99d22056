// Suffix array by induced sorting (SA-IS, Nong/Zhang/Chan 2009), written for this repo.
//
// Role: host-side index PRODUCER only.  The reference builds its suffix array with the
// third-party jsuffixarrays DivSufSort (indices/src/main/java/com/dynatrace/fm/FmIndex.java:332-341),
// which is not vendored in /root/reference.  The suffix array of a text that ends in a unique
// smallest sentinel is unique, so any correct algorithm yields the identical array.
//
// Requirements: s[n-1] is the unique smallest symbol (the FM-index sentinel, code 0), symbols are
// in [0, K).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace fmhost {

namespace sais_detail {

template <typename T>
static void bucket_bounds(const T* s, int32_t* bkt, int32_t n, int32_t K, bool end) {
    for (int32_t i = 0; i < K; ++i) bkt[i] = 0;
    for (int32_t i = 0; i < n; ++i) bkt[s[i]]++;
    int32_t sum = 0;
    for (int32_t i = 0; i < K; ++i) {
        sum += bkt[i];
        bkt[i] = end ? sum : sum - bkt[i];
    }
}

struct TypeBits {
    std::vector<uint8_t> b;
    explicit TypeBits(int32_t n) : b((size_t)n / 8 + 1, 0) {}
    inline bool get(int32_t i) const { return (b[i >> 3] >> (i & 7)) & 1; }  // 1 = S-type
    inline void set(int32_t i, bool v) {
        if (v) b[i >> 3] |= (uint8_t)(1u << (i & 7));
        else b[i >> 3] &= (uint8_t)~(1u << (i & 7));
    }
    inline bool is_lms(int32_t i) const { return i > 0 && get(i) && !get(i - 1); }
};

template <typename T>
static void induce_l(const TypeBits& t, int32_t* SA, const T* s, int32_t* bkt, int32_t n, int32_t K) {
    bucket_bounds(s, bkt, n, K, false);
    for (int32_t i = 0; i < n; ++i) {
        int32_t j = SA[i] - 1;
        if (j >= 0 && !t.get(j)) SA[bkt[s[j]]++] = j;
    }
}

template <typename T>
static void induce_s(const TypeBits& t, int32_t* SA, const T* s, int32_t* bkt, int32_t n, int32_t K) {
    bucket_bounds(s, bkt, n, K, true);
    for (int32_t i = n - 1; i >= 0; --i) {
        int32_t j = SA[i] - 1;
        if (j >= 0 && t.get(j)) SA[--bkt[s[j]]] = j;
    }
}

template <typename T>
static void sais_rec(const T* s, int32_t* SA, int32_t n, int32_t K) {
    if (n == 1) {
        SA[0] = 0;
        return;
    }
    TypeBits t(n);
    t.set(n - 1, true);
    t.set(n - 2, false);
    for (int32_t i = n - 3; i >= 0; --i)
        t.set(i, s[i] < s[i + 1] || (s[i] == s[i + 1] && t.get(i + 1)));

    std::vector<int32_t> bkt_store((size_t)K);
    int32_t* bkt = bkt_store.data();

    // stage 1: sort LMS substrings
    bucket_bounds(s, bkt, n, K, true);
    for (int32_t i = 0; i < n; ++i) SA[i] = -1;
    for (int32_t i = 1; i < n; ++i)
        if (t.is_lms(i)) SA[--bkt[s[i]]] = i;
    induce_l(t, SA, s, bkt, n, K);
    induce_s(t, SA, s, bkt, n, K);

    // compact sorted LMS substrings into SA[0, n1)
    int32_t n1 = 0;
    for (int32_t i = 0; i < n; ++i)
        if (t.is_lms(SA[i])) SA[n1++] = SA[i];
    for (int32_t i = n1; i < n; ++i) SA[i] = -1;

    // name LMS substrings
    int32_t name = 0, prev = -1;
    for (int32_t i = 0; i < n1; ++i) {
        int32_t pos = SA[i];
        bool diff = false;
        for (int32_t d = 0; d < n; ++d) {
            if (prev == -1 || s[pos + d] != s[prev + d] || t.get(pos + d) != t.get(prev + d)) {
                diff = true;
                break;
            } else if (d > 0 && (t.is_lms(pos + d) || t.is_lms(prev + d))) {
                break;
            }
        }
        if (diff) {
            ++name;
            prev = pos;
        }
        SA[n1 + (pos >> 1)] = name - 1;
    }
    for (int32_t i = n - 1, j = n - 1; i >= n1; --i)
        if (SA[i] >= 0) SA[j--] = SA[i];

    // stage 2: solve the reduced problem
    int32_t* SA1 = SA;
    int32_t* s1 = SA + n - n1;
    if (name < n1) {
        sais_rec<int32_t>(s1, SA1, n1, name);
    } else {
        for (int32_t i = 0; i < n1; ++i) SA1[s1[i]] = i;
    }

    // stage 3: induce the result
    bucket_bounds(s, bkt, n, K, true);
    for (int32_t i = 1, j = 0; i < n; ++i)
        if (t.is_lms(i)) s1[j++] = i;
    for (int32_t i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
    for (int32_t i = n1; i < n; ++i) SA[i] = -1;
    for (int32_t i = n1 - 1; i >= 0; --i) {
        int32_t j = SA[i];
        SA[i] = -1;
        SA[--bkt[s[j]]] = j;
    }
    induce_l(t, SA, s, bkt, n, K);
    induce_s(t, SA, s, bkt, n, K);
}

}  // namespace sais_detail

// s: n symbols in [0,K), s[n-1] unique smallest.  SA: n entries out.
template <typename T>
inline void suffix_array(const T* s, int32_t* SA, int32_t n, int32_t K) {
    sais_detail::sais_rec<T>(s, SA, n, K);
}

}  // namespace fmhost
