// TEST INFRASTRUCTURE (tests/test_sort_replay.py): the product's std::sort replay (csrc/std_sort_replay.cuh, the code
// the device runs) compiled for the host and compared with libstdc++'s std::sort itself, element for element, on the
// two sorts of the path: match list by distance DESCENDING (match_features.cpp:100-101) and PROSAC pool by quality
// ASCENDING (ransac.cpp:83-90). Inputs: random keys with many ties (Hamming distances), all equal, sorted, reversed,
// organ pipe, and inputs built with McIlroy's adversary ("A killer adversary for quicksort", 1999) against std::sort
// itself, which exhaust the depth limit and force the heap-sort fallback.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

static long g_heap_sorts = 0;
#define OCB_SORT_REPLAY_ON_HEAP_SORT (++g_heap_sorts)
#include "../opencalibration_b200/csrc/std_sort_replay.cuh"

struct Rec
{
    uint32_t a, k, d; // like ocb_match: the comparator looks at d only
};

template <bool DESC> static bool check(const std::vector<uint32_t> &keys, const char *what)
{
    const size_t n = keys.size();
    std::vector<Rec> ref(n);
    std::vector<uint64_t> v(n);
    for (size_t i = 0; i < n; i++)
    {
        ref[i] = Rec{(uint32_t)i, (uint32_t)(i * 7), keys[i]};
        v[i] = ((uint64_t)keys[i] << 32) | (uint32_t)i;
    }
    if (DESC)
        std::sort(ref.begin(), ref.end(), [](const Rec &x, const Rec &y) { return x.d > y.d; });
    else
        std::sort(ref.begin(), ref.end(), [](const Rec &x, const Rec &y) { return x.d < y.d; });
    ocb::sort_replay::std_sort(v.data(), (long)n, ocb::sort_replay::KeyOrder<DESC>());
    for (size_t i = 0; i < n; i++)
        if ((uint32_t)v[i] != ref[i].a)
        {
            std::printf("MISMATCH %s n=%zu at %zu: replay %u, std::sort %u\n", what, n, i, (uint32_t)v[i], ref[i].a);
            return false;
        }
    return true;
}

// McIlroy's adversary run against std::sort -> an input on which std::sort's quicksort phase degenerates
static std::vector<uint32_t> killer(size_t n)
{
    std::vector<int> val(n, (int)n - 1), ptr(n);
    const int gas = (int)n - 1;
    int nsolid = 0, candidate = 0;
    for (size_t i = 0; i < n; i++)
        ptr[i] = (int)i;
    std::sort(ptr.begin(), ptr.end(), [&](int x, int y) {
        if (val[x] == gas && val[y] == gas)
        {
            if (x == candidate)
                val[x] = nsolid++;
            else
                val[y] = nsolid++;
        }
        if (val[x] == gas)
            candidate = x;
        else if (val[y] == gas)
            candidate = y;
        return val[x] < val[y];
    });
    return std::vector<uint32_t>(val.begin(), val.end());
}

int main()
{
    std::mt19937_64 g(12345);
    long cases = 0;
    bool ok = true;
    auto both = [&](const std::vector<uint32_t> &k, const char *what) {
        ok = check<true>(k, what) && ok;
        ok = check<false>(k, what) && ok;
        cases += 2;
    };
    for (int trial = 0; trial < 6000 && ok; trial++)
    {
        const size_t n = trial < 200 ? (size_t)trial : (size_t)(g() % 5000);
        const uint32_t spread = (uint32_t)(1 + g() % (trial % 3 == 0 ? 8 : 487));
        std::vector<uint32_t> k(n);
        for (auto &x : k)
            x = (uint32_t)(g() % spread);
        both(k, "random");
        if (trial % 10 == 0)
        {
            std::sort(k.begin(), k.end());
            both(k, "sorted");
            std::reverse(k.begin(), k.end());
            both(k, "reversed");
            for (size_t i = 0; i < n; i++)
                k[i] = (uint32_t)std::min(i, n - 1 - i);
            both(k, "organ pipe");
            std::fill(k.begin(), k.end(), 7u);
            both(k, "all equal");
        }
    }
    const long before = g_heap_sorts;
    for (size_t n : {17u, 64u, 500u, 2000u, 8192u, 30000u})
    {
        std::vector<uint32_t> k = killer(n);
        ok = check<false>(k, "killer") && ok; // built against the ascending comparator
        std::vector<uint32_t> flipped(k);
        for (auto &x : flipped)
            x = (uint32_t)n - x;
        ok = check<true>(flipped, "killer, descending") && ok;
        cases += 2;
    }
    const long fallbacks = g_heap_sorts - before;
    std::printf("%s: %ld cases, heap-sort fallback taken %ld times on the adversarial inputs\n", ok ? "OK" : "FAILED", cases,
                fallbacks);
    return ok && fallbacks > 0 ? 0 : 1;
}
