// libstdc++'s std::sort, replayed step for step so that it can run on the device.
//
// The reference orders its match list with std::sort and a comparator on the distance only
// (src/match/match_features.cpp:100-101) and its PROSAC pool with std::sort on the quality only
// (src/model_inliers/ransac.cpp:83-90). std::sort is not stable: the order among equal keys is whatever the library's
// introsort produces from the input order, and everything downstream (PROSAC samples, hence the RANSAC result) depends
// on it. The sequence of comparisons and moves of that algorithm depends only on the comparator's answers, so running
// the SAME algorithm on (key, original position) words yields the reference's permutation. This header restates
// GCC 13's bits/stl_algo.h / bits/stl_heap.h:
//   __sort            -> __introsort_loop(first, last, 2 * __lg(n)) + __final_insertion_sort      (:1937-1950)
//   __introsort_loop  -> while (n > 16): depth exhausted ? __partial_sort (heap sort) : __unguarded_partition_pivot;
//                        recurse on the right part, loop on the left                                (:1908-1931)
//   __unguarded_partition_pivot -> __move_median_to_first(first, first + 1, mid, last - 1) + __unguarded_partition
//   __final_insertion_sort -> __insertion_sort on the first 16, __unguarded_insertion_sort on the rest (:1813-1830)
//   __partial_sort(first, last, last) -> __make_heap + __sort_heap (__adjust_heap / __push_heap, stl_heap.h)
// The two parts a partition step creates are disjoint, so the order in which they are finished does not matter; the
// recursion is an explicit stack here. tests/test_sort_replay.py compiles this header for the host and checks it
// against std::sort itself on random, tied, sorted, reversed and median-of-three-killer inputs.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define OCB_HD __host__ __device__ __forceinline__
#else
#define OCB_HD inline
#endif

namespace ocb
{
namespace sort_replay
{
// Element = (key << 32) | payload. Only the key takes part in comparisons, like the reference's comparators.
template <bool DESCENDING> struct KeyOrder
{
    OCB_HD bool operator()(uint64_t a, uint64_t b) const
    {
        const uint32_t ka = (uint32_t)(a >> 32), kb = (uint32_t)(b >> 32);
        return DESCENDING ? ka > kb : ka < kb;
    }
};

template <typename Comp> OCB_HD void push_heap(uint64_t *first, long hole, long top, uint64_t value, Comp comp)
{
    long parent = (hole - 1) / 2;
    while (hole > top && comp(first[parent], value))
    {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

template <typename Comp> OCB_HD void adjust_heap(uint64_t *first, long hole, long len, uint64_t value, Comp comp)
{
    const long top = hole;
    long second = hole;
    while (second < (len - 1) / 2)
    {
        second = 2 * (second + 1);
        if (comp(first[second], first[second - 1]))
            second--;
        first[hole] = first[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2)
    {
        second = 2 * (second + 1);
        first[hole] = first[second - 1];
        hole = second - 1;
    }
    push_heap(first, hole, top, value, comp);
}

// __partial_sort(first, last, last): heap sort of the whole range
template <typename Comp> OCB_HD void heap_sort(uint64_t *first, long len, Comp comp)
{
    if (len >= 2) // __make_heap
    {
        long parent = (len - 2) / 2;
        for (;;)
        {
            const uint64_t value = first[parent];
            adjust_heap(first, parent, len, value, comp);
            if (parent == 0)
                break;
            parent--;
        }
    }
    long last = len; // __sort_heap
    while (last > 1)
    {
        --last;
        const uint64_t value = first[last]; // __pop_heap(first, last, last)
        first[last] = first[0];
        adjust_heap(first, 0, last, value, comp);
    }
}

template <typename Comp> OCB_HD void swap_at(uint64_t *v, long a, long b)
{
    const uint64_t t = v[a];
    v[a] = v[b];
    v[b] = t;
}

template <typename Comp> OCB_HD void move_median_to_first(uint64_t *v, long result, long a, long b, long c, Comp comp)
{
    if (comp(v[a], v[b]))
    {
        if (comp(v[b], v[c]))
            swap_at<Comp>(v, result, b);
        else if (comp(v[a], v[c]))
            swap_at<Comp>(v, result, c);
        else
            swap_at<Comp>(v, result, a);
    }
    else if (comp(v[a], v[c]))
        swap_at<Comp>(v, result, a);
    else if (comp(v[b], v[c]))
        swap_at<Comp>(v, result, c);
    else
        swap_at<Comp>(v, result, b);
}

template <typename Comp> OCB_HD long unguarded_partition(uint64_t *v, long first, long last, long pivot, Comp comp)
{
    for (;;)
    {
        while (comp(v[first], v[pivot]))
            ++first;
        --last;
        while (comp(v[pivot], v[last]))
            --last;
        if (!(first < last))
            return first;
        swap_at<Comp>(v, first, last);
        ++first;
    }
}

template <typename Comp> OCB_HD void unguarded_linear_insert(uint64_t *v, long last, Comp comp)
{
    const uint64_t val = v[last];
    long next = last - 1;
    while (comp(val, v[next]))
    {
        v[last] = v[next];
        last = next;
        --next;
    }
    v[last] = val;
}

template <typename Comp> OCB_HD void insertion_sort(uint64_t *v, long first, long last, Comp comp)
{
    if (first == last)
        return;
    for (long i = first + 1; i != last; ++i)
    {
        if (comp(v[i], v[first]))
        {
            const uint64_t val = v[i];
            for (long k = i; k > first; --k) // move_backward(first, i, i + 1)
                v[k] = v[k - 1];
            v[first] = val;
        }
        else
            unguarded_linear_insert(v, i, comp);
    }
}

constexpr long THRESHOLD = 16; // _S_threshold

OCB_HD long floor_log2(long n) // std::__lg
{
    long k = 0;
    while (n > 1)
    {
        n >>= 1;
        k++;
    }
    return k;
}

// std::sort(v, v + n, comp)
template <typename Comp> OCB_HD void std_sort(uint64_t *v, long n, Comp comp)
{
    if (n <= 0)
        return;
    // __introsort_loop with the recursion on an explicit stack (depth <= 2 * lg(n) + 1 <= 129 entries for n < 2^63)
    long stack_lo[132], stack_hi[132], stack_depth[132];
    int sp = 0;
    stack_lo[0] = 0, stack_hi[0] = n, stack_depth[0] = 2 * floor_log2(n);
    sp = 1;
    while (sp > 0)
    {
        --sp;
        long first = stack_lo[sp], last = stack_hi[sp], depth = stack_depth[sp];
        while (last - first > THRESHOLD)
        {
            if (depth == 0)
            {
#ifdef OCB_SORT_REPLAY_ON_HEAP_SORT
                OCB_SORT_REPLAY_ON_HEAP_SORT; // test hook: tells the checker that the fallback really ran
#endif
                heap_sort(v + first, last - first, comp);
                break;
            }
            --depth;
            const long mid = first + (last - first) / 2;
            move_median_to_first(v, first, first + 1, mid, last - 1, comp);
            const long cut = unguarded_partition(v, first + 1, last, first, comp);
            stack_lo[sp] = cut, stack_hi[sp] = last, stack_depth[sp] = depth; // __introsort_loop(cut, last, depth)
            sp++;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > THRESHOLD)
    {
        insertion_sort(v, 0, THRESHOLD, comp);
        for (long i = THRESHOLD; i != n; ++i)
            unguarded_linear_insert(v, i, comp);
    }
    else
        insertion_sort(v, 0, n, comp);
}
} // namespace sort_replay
} // namespace ocb
