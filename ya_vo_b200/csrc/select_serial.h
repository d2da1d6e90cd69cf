// select_serial.h — single-thread pieces of the exact std::sort replay.
//
// The reference orders FAST candidates with libstdc++'s std::sort and a comparator on the Harris
// response only (reference src/FastDetector.cc:343-345).  std::sort is unstable, so the order of
// candidates with equal responses is whatever introsort's partition history produces.  To return
// the reference's keypoint order exactly, the CUDA select kernel replays that algorithm
// (bits/stl_algo.h: __introsort_loop / __unguarded_partition_pivot / __move_median_to_first /
// __final_insertion_sort; bits/stl_heap.h for the depth-limit fallback) on the candidate list in
// scan order, restricted to the ranges that can reach the first K outputs.
//
// This header holds the parts a single thread runs (small ranges, heapsort fallback).  It compiles
// for the device and, for the CPU model in tests/, for the host.
#ifndef YAVO_SELECT_SERIAL_H
#define YAVO_SELECT_SERIAL_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define YAVO_HD __host__ __device__ __forceinline__
#else
#define YAVO_HD inline
#endif

// One candidate: high 32 bits = IEEE-754 bits of the float score, low 32 bits = payload (row<<16 | col).
typedef unsigned long long yavo_ent;

YAVO_HD float yavo_ent_score(yavo_ent e) {
#ifdef __CUDA_ARCH__
    return __uint_as_float((unsigned)(e >> 32));
#else
    uint32_t u = (uint32_t)(e >> 32);
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
YAVO_HD yavo_ent yavo_make_ent(float score, uint32_t payload) {
#ifdef __CUDA_ARCH__
    return ((yavo_ent)__float_as_uint(score) << 32) | payload;
#else
    uint32_t u;
    memcpy(&u, &score, 4);
    return ((yavo_ent)u << 32) | payload;
#endif
}
// the reference comparator: a sorts before b iff a.cornerResponse > b.cornerResponse
YAVO_HD bool yavo_before(yavo_ent a, yavo_ent b) { return yavo_ent_score(a) > yavo_ent_score(b); }

#define YAVO_SORT_THRESHOLD 16  // libstdc++ _S_threshold

// __move_median_to_first(result=first, a=first+1, b=mid, c=last-1)
template <typename P>
YAVO_HD void yavo_median_to_first(P A, int first, int last) {
    int a = first + 1, b = first + (last - first) / 2, c = last - 1;
    yavo_ent va = A[a], vb = A[b], vc = A[c];
    int m;
    if (yavo_before(va, vb)) {
        if (yavo_before(vb, vc)) m = b;
        else if (yavo_before(va, vc)) m = c;
        else m = a;
    } else if (yavo_before(va, vc)) m = a;
    else if (yavo_before(vb, vc)) m = c;
    else m = b;
    yavo_ent t = A[first];
    A[first] = A[m];
    A[m] = t;
}

// __unguarded_partition(first+1, last, pivot=first): returns the cut
template <typename P>
YAVO_HD int yavo_serial_partition(P A, int first, int last) {
    yavo_median_to_first(A, first, last);
    const yavo_ent piv = A[first];
    int f = first + 1, l = last;
    for (;;) {
        while (yavo_before(A[f], piv)) ++f;
        --l;
        while (yavo_before(piv, A[l])) --l;
        if (!(f < l)) return f;
        yavo_ent t = A[f];
        A[f] = A[l];
        A[l] = t;
        ++f;
    }
}

// __insertion_sort / __unguarded_linear_insert: a stable insertion sort of [first,last)
template <typename P>
YAVO_HD void yavo_serial_insertion(P A, int first, int last) {
    for (int i = first + 1; i < last; i++) {
        yavo_ent v = A[i];
        int j = i;
        while (j > first && yavo_before(v, A[j - 1])) {
            A[j] = A[j - 1];
            j--;
        }
        A[j] = v;
    }
}

// bits/stl_heap.h __push_heap
template <typename P>
YAVO_HD void yavo_push_heap(P A, int base, int hole, int top, yavo_ent value) {
    int parent = (hole - 1) / 2;
    while (hole > top && yavo_before(A[base + parent], value)) {
        A[base + hole] = A[base + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    A[base + hole] = value;
}
// bits/stl_heap.h __adjust_heap
template <typename P>
YAVO_HD void yavo_adjust_heap(P A, int base, int hole, int len, yavo_ent value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (yavo_before(A[base + child], A[base + child - 1])) child--;
        A[base + hole] = A[base + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        A[base + hole] = A[base + child - 1];
        hole = child - 1;
    }
    yavo_push_heap(A, base, hole, top, value);
}
// __partial_sort(first, last, last) == __make_heap + __sort_heap
template <typename P>
YAVO_HD void yavo_serial_heapsort(P A, int first, int last) {
    int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        for (;;) {
            yavo_ent v = A[first + parent];
            yavo_adjust_heap(A, first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    while (len > 1) {
        --len;  // __pop_heap(first, last-1, last-1)
        yavo_ent v = A[first + len];
        A[first + len] = A[first];
        yavo_adjust_heap(A, first, 0, len, v);
    }
}

// __introsort_loop on [first,last) with the given remaining depth, followed by the part of
// __final_insertion_sort that touches this range.  Ranges that start at or beyond k cannot
// influence outputs [0,k) and are skipped.
template <typename P>
YAVO_HD void yavo_serial_introsort(P A, int first, int last, int depth, int k) {
    int stk_f[64], stk_l[64], stk_d[64];
    int sp = 0;
    for (;;) {
        bool finished_leaf = false;
        while (last - first > YAVO_SORT_THRESHOLD) {
            if (first >= k) { finished_leaf = true; break; }
            if (depth == 0) {
                yavo_serial_heapsort(A, first, last);
                finished_leaf = true;
                break;
            }
            --depth;
            int cut = yavo_serial_partition(A, first, last);
            // right part [cut,last): defer (order between disjoint ranges is immaterial)
            if (cut < k) {
                if (last - cut > YAVO_SORT_THRESHOLD && sp < 64) {
                    stk_f[sp] = cut; stk_l[sp] = last; stk_d[sp] = depth; sp++;
                } else if (last - cut > YAVO_SORT_THRESHOLD) {
                    yavo_serial_heapsort(A, cut, last);  // unreachable: depth <= 2*lg(n) < 64
                } else {
                    yavo_serial_insertion(A, cut, last);
                }
            }
            last = cut;
        }
        if (!finished_leaf && first < k) yavo_serial_insertion(A, first, last);
        if (sp == 0) break;
        sp--;
        first = stk_f[sp]; last = stk_l[sp]; depth = stk_d[sp];
    }
}

#endif  // YAVO_SELECT_SERIAL_H
