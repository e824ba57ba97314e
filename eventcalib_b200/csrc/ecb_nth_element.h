// std::nth_element as libstdc++ (g++ 13, bits/stl_algo.h, bits/stl_heap.h) executes it, restated over an index
// range so that the element that ends up in slot `nth` — including which one of several EQUAL keys — is the one the
// reference gets at CirclesEventFrame.cpp:140-147 (the cluster "centre" = member with the median norm).
//
//   nth_element(first, nth, last, comp) -> __introselect(first, nth, last, 2 * lg(last - first), comp)        [external]
//     while (last - first > 3):  depth limit hit -> __heap_select(first, nth + 1, last) ; iter_swap(first, nth) ; return
//                                cut = __unguarded_partition_pivot(first, last)   (median of first+1, mid, last-1 -> first)
//                                cut <= nth ? first = cut : last = cut
//     __insertion_sort(first, last)
//
// Host + device, no dependencies; `Less(a, b)` compares two stored VALUES (here: pids by squared norm).
// tests/test_nth_element.py compiles this header with g++ and checks it against the real std::nth_element.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ECB_HD __host__ __device__ __forceinline__
#else
#define ECB_HD inline
#endif

namespace ecb_nth {

template <typename V>
ECB_HD void swap_at(V *a, long i, long j) {
    V t = a[i];
    a[i] = a[j];
    a[j] = t;
}

// __move_median_to_first(result, a, b, c)
template <typename V, typename Less>
ECB_HD void move_median_to_first(V *x, long result, long a, long b, long c, Less less) {
    if (less(x[a], x[b])) {
        if (less(x[b], x[c])) swap_at(x, result, b);
        else if (less(x[a], x[c])) swap_at(x, result, c);
        else swap_at(x, result, a);
    } else if (less(x[a], x[c])) swap_at(x, result, a);
    else if (less(x[b], x[c])) swap_at(x, result, c);
    else swap_at(x, result, b);
}

// __unguarded_partition(first, last, pivot)
template <typename V, typename Less>
ECB_HD long unguarded_partition(V *x, long first, long last, long pivot, Less less) {
    for (;;) {
        while (less(x[first], x[pivot])) ++first;
        --last;
        while (less(x[pivot], x[last])) --last;
        if (!(first < last)) return first;
        swap_at(x, first, last);
        ++first;
    }
}

// __push_heap(first, holeIndex, topIndex, value)
template <typename V, typename Less>
ECB_HD void push_heap_(V *x, long first, long hole, long top, V value, Less less) {
    long parent = (hole - 1) / 2;
    while (hole > top && less(x[first + parent], value)) {
        x[first + hole] = x[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    x[first + hole] = value;
}

// __adjust_heap(first, holeIndex, len, value)
template <typename V, typename Less>
ECB_HD void adjust_heap(V *x, long first, long hole, long len, V value, Less less) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(x[first + child], x[first + (child - 1)])) child--;
        x[first + hole] = x[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        x[first + hole] = x[first + (child - 1)];
        hole = child - 1;
    }
    push_heap_(x, first, hole, top, value, less);
}

// __heap_select(first, middle, last): __make_heap(first, middle) then __pop_heap for every smaller element
template <typename V, typename Less>
ECB_HD void heap_select(V *x, long first, long middle, long last, Less less) {
    const long len = middle - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            V value = x[first + parent];
            adjust_heap(x, first, parent, len, value, less);
            if (parent == 0) break;
            parent--;
        }
    }
    for (long i = middle; i < last; ++i)
        if (less(x[i], x[first])) {  // __pop_heap(first, middle, i)
            V value = x[i];
            x[i] = x[first];
            adjust_heap(x, first, 0, len, value, less);
        }
}

// __insertion_sort(first, last)
template <typename V, typename Less>
ECB_HD void insertion_sort(V *x, long first, long last, Less less) {
    if (first == last) return;
    for (long i = first + 1; i != last; ++i) {
        if (less(x[i], x[first])) {
            V val = x[i];
            for (long j = i; j > first; --j) x[j] = x[j - 1];  // move_backward(first, i, i + 1)
            x[first] = val;
        } else {  // __unguarded_linear_insert(i)
            V val = x[i];
            long last_ = i, next = i - 1;
            while (less(val, x[next])) {
                x[last_] = x[next];
                last_ = next;
                --next;
            }
            x[last_] = val;
        }
    }
}

template <typename V, typename Less>
ECB_HD void nth_element(V *x, long n, long nth, Less less) {
    long first = 0, last = n;
    if (first == last || nth == last) return;
    long depth_limit = 0;
    for (long k = n; k > 1; k >>= 1) ++depth_limit;  // std::__lg(n)
    depth_limit *= 2;
    while (last - first > 3) {
        if (depth_limit == 0) {
            heap_select(x, first, nth + 1, last, less);
            swap_at(x, first, nth);
            return;
        }
        --depth_limit;
        const long mid = first + (last - first) / 2;
        move_median_to_first(x, first, first + 1, mid, last - 1, less);
        const long cut = unguarded_partition(x, first + 1, last, first, less);
        if (cut <= nth) first = cut;
        else last = cut;
    }
    insertion_sort(x, first, last, less);
}

}  // namespace ecb_nth
