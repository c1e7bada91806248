// Suffix array construction by induced sorting (SA-IS, Nong/Zhang/Chan 2009).
//
// FIXTURE CODE: the reference delegates suffix sorting to external binaries (Big-BWT / gsacak,
// /root/reference/src/spumoni.cpp:442-472) that are not available in this image; tests and the
// benchmark need indexes, so this header provides an in-repo suffix sorter.  It is not on the
// query hot path.
//
// Requirements: s[n-1] must be the unique, smallest symbol (a sentinel).  Idx is a signed integer
// type (int32_t for n < 2^31, int64_t above).
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace sais_detail {

struct TypeBits {
    std::vector<uint64_t> w;
    explicit TypeBits(size_t n) : w((n + 63) / 64, 0) {}
    inline bool get(size_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    inline void set(size_t i, bool b) {
        if (b) w[i >> 6] |= (uint64_t(1) << (i & 63));
        else w[i >> 6] &= ~(uint64_t(1) << (i & 63));
    }
};

template <typename Sym, typename Idx>
static void get_buckets(const Sym* s, Idx* bkt, Idx n, Idx K, bool end) {
    for (Idx i = 0; i <= K; i++) bkt[i] = 0;
    for (Idx i = 0; i < n; i++) bkt[(size_t)s[i]]++;
    Idx sum = 0;
    for (Idx i = 0; i <= K; i++) {
        sum += bkt[i];
        bkt[i] = end ? sum : sum - bkt[i];
    }
}

template <typename Sym, typename Idx>
static void induce_L(const TypeBits& t, Idx* SA, const Sym* s, Idx* bkt, Idx n, Idx K) {
    get_buckets(s, bkt, n, K, false);
    for (Idx i = 0; i < n; i++) {
        Idx j = SA[i] - 1;
        if (SA[i] > 0 && !t.get((size_t)j)) SA[bkt[(size_t)s[j]]++] = j;
    }
}

template <typename Sym, typename Idx>
static void induce_S(const TypeBits& t, Idx* SA, const Sym* s, Idx* bkt, Idx n, Idx K) {
    get_buckets(s, bkt, n, K, true);
    for (Idx i = n - 1; i >= 0; i--) {
        Idx j = SA[i] - 1;
        if (SA[i] > 0 && t.get((size_t)j)) SA[--bkt[(size_t)s[j]]] = j;
    }
}

template <typename Sym, typename Idx>
static void sais_rec(const Sym* s, Idx* SA, Idx n, Idx K) {
    TypeBits t((size_t)n);
    auto isLMS = [&](Idx i) { return i > 0 && t.get((size_t)i) && !t.get((size_t)i - 1); };
    // classify: S-type = 1, L-type = 0
    t.set((size_t)n - 1, true);
    if (n >= 2) t.set((size_t)n - 2, false);
    for (Idx i = n - 3; i >= 0; i--)
        t.set((size_t)i, (s[i] < s[i + 1] || (s[i] == s[i + 1] && t.get((size_t)i + 1))));

    std::vector<Idx> bkt((size_t)K + 1);
    // stage 1: sort all LMS substrings
    get_buckets(s, bkt.data(), n, K, true);
    for (Idx i = 0; i < n; i++) SA[i] = -1;
    for (Idx i = 1; i < n; i++)
        if (isLMS(i)) SA[--bkt[(size_t)s[i]]] = i;
    induce_L(t, SA, s, bkt.data(), n, K);
    induce_S(t, SA, s, bkt.data(), n, K);

    // compact sorted LMS substrings into SA[0..n1)
    Idx n1 = 0;
    for (Idx i = 0; i < n; i++)
        if (isLMS(SA[i])) SA[n1++] = SA[i];
    for (Idx i = n1; i < n; i++) SA[i] = -1;
    // name them
    Idx name = 0, prev = -1;
    for (Idx i = 0; i < n1; i++) {
        Idx pos = SA[i];
        bool diff = false;
        for (Idx d = 0; d < n; d++) {
            if (prev == -1 || s[pos + d] != s[prev + d] || t.get((size_t)(pos + d)) != t.get((size_t)(prev + d))) {
                diff = true;
                break;
            } else if (d > 0 && (isLMS(pos + d) || isLMS(prev + d)))
                break;
        }
        if (diff) {
            name++;
            prev = pos;
        }
        SA[n1 + pos / 2] = name - 1;
    }
    for (Idx i = n - 1, j = n - 1; i >= n1; i--)
        if (SA[i] >= 0) SA[j--] = SA[i];

    // stage 2: solve the reduced problem
    Idx* SA1 = SA;
    Idx* s1 = SA + n - n1;
    if (name < n1)
        sais_rec<Idx, Idx>(s1, SA1, n1, name - 1);
    else
        for (Idx i = 0; i < n1; i++) SA1[s1[i]] = i;

    // stage 3: induce the final SA
    get_buckets(s, bkt.data(), n, K, true);
    for (Idx i = 1, j = 0; i < n; i++)
        if (isLMS(i)) s1[j++] = i;
    for (Idx i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
    for (Idx i = n1; i < n; i++) SA[i] = -1;
    for (Idx i = n1 - 1; i >= 0; i--) {
        Idx j = SA[i];
        SA[i] = -1;
        SA[--bkt[(size_t)s[j]]] = j;
    }
    induce_L(t, SA, s, bkt.data(), n, K);
    induce_S(t, SA, s, bkt.data(), n, K);
}

}  // namespace sais_detail

// s: n symbols, s[n-1] unique smallest; SA: n entries out.
template <typename Idx>
inline void suffix_array_sais(const uint8_t* s, Idx* SA, Idx n) {
    if (n == 1) {
        SA[0] = 0;
        return;
    }
    sais_detail::sais_rec<uint8_t, Idx>(s, SA, n, (Idx)255);
}
