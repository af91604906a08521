/*
 * Oracle port — the candidate order of the reference. TEST INFRASTRUCTURE ONLY (see mzo.h).
 *
 * ZeroActor::calculateAlphaZeroActionPolicy / calculateMuZeroActionPolicy (actor/zero_actor.cpp:215-245) order the
 * candidates with std::sort(begin, end, lhs.policy_ > rhs.policy_). std::sort is unstable: the order of candidates with
 * EQUAL priors is whatever the library's algorithm leaves, and it is observable (child order decides PUCT ties and the
 * record's P[...] tag). The algorithm is a third-party dependency that is not under /root/reference: libstdc++
 * (GCC; the reference's docker image and this container both use it), bits/stl_algo.h std::__sort — introsort:
 * median-of-three quicksort down to ranges of 16 with a depth limit of 2*floor(log2 n) (heapsort beyond it), then one
 * insertion-sort pass — unchanged since GCC 4.9. It is restated here operation by operation (stl_algo.h:84-103,1792-1950,
 * stl_heap.h:135-264,340-432 of GCC 13) on candidates given in ascending action id, the order the reference pushes them.
 */
#include "mzo.h"

typedef struct {
    float p, l;
    int a;
} cand;

static int gt(const cand* x, const cand* y) { return x->p > y->p; } /* the comparator: lhs.policy_ > rhs.policy_ */

static void swap_c(cand* x, cand* y)
{
    cand t = *x;
    *x = *y;
    *y = t;
}

static void move_median_to_first(cand* result, cand* a, cand* b, cand* c)
{
    if (gt(a, b)) {
        if (gt(b, c)) {
            swap_c(result, b);
        } else if (gt(a, c)) {
            swap_c(result, c);
        } else {
            swap_c(result, a);
        }
    } else if (gt(a, c)) {
        swap_c(result, a);
    } else if (gt(b, c)) {
        swap_c(result, c);
    } else {
        swap_c(result, b);
    }
}

static cand* unguarded_partition(cand* first, cand* last, cand* pivot)
{
    for (;;) {
        while (gt(first, pivot)) { ++first; }
        --last;
        while (gt(pivot, last)) { --last; }
        if (!(first < last)) { return first; }
        swap_c(first, last);
        ++first;
    }
}

static void push_heap(cand* first, long hole, long top, cand value)
{
    long parent = (hole - 1) / 2;
    while (hole > top && gt(first + parent, &value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void adjust_heap(cand* first, long hole, long len, cand value)
{
    const long top = hole;
    long second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (gt(first + second, first + (second - 1))) { second--; }
        first[hole] = first[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        first[hole] = first[second - 1];
        hole = second - 1;
    }
    push_heap(first, hole, top, value);
}

static void heap_sort(cand* first, cand* last) /* __partial_sort(first, last, last): make_heap + sort_heap */
{
    const long len = last - first;
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            cand value = first[parent];
            adjust_heap(first, parent, len, value);
            if (parent == 0) { break; }
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        cand value = *last;
        *last = *first;
        adjust_heap(first, 0, last - first, value);
    }
}

static void introsort_loop(cand* first, cand* last, int depth_limit)
{
    while (last - first > 16) {
        if (depth_limit == 0) {
            heap_sort(first, last);
            return;
        }
        --depth_limit;
        cand* mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        cand* cut = unguarded_partition(first + 1, last, first);
        introsort_loop(cut, last, depth_limit);
        last = cut;
    }
}

static void unguarded_linear_insert(cand* last)
{
    cand val = *last;
    cand* next = last - 1;
    while (gt(&val, next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

static void insertion_sort(cand* first, cand* last)
{
    if (first == last) { return; }
    for (cand* i = first + 1; i != last; ++i) {
        if (gt(i, first)) {
            cand val = *i;
            for (cand* j = i; j != first; --j) { *j = *(j - 1); }
            *first = val;
        } else {
            unguarded_linear_insert(i);
        }
    }
}

/* candidates (action ids ascending on entry) -> the order std::sort leaves them in */
void mzo_std_sort_candidates(int n, int32_t* action, float* policy, float* logit)
{
    cand c[MZO_MAX_ACTIONS];
    for (int i = 0; i < n; ++i) { c[i].a = action[i], c[i].p = policy[i], c[i].l = logit[i]; }
    if (n > 0) {
        int lg = 0;
        for (int m = n; m > 1; m >>= 1) { ++lg; }
        introsort_loop(c, c + n, lg * 2);
        if (n > 16) {
            insertion_sort(c, c + 16);
            for (cand* i = c + 16; i != c + n; ++i) { unguarded_linear_insert(i); }
        } else {
            insertion_sort(c, c + n);
        }
    }
    for (int i = 0; i < n; ++i) { action[i] = c[i].a, policy[i] = c[i].p, logit[i] = c[i].l; }
}
