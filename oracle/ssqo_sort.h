/*
 * ssqo_sort.h — ORACLE (test infrastructure): an UNSTABLE introsort whose sequence of comparisons and
 * swaps restates klib's ks_introsort (median-of-3 quicksort, partitions <=16 left to one final
 * insertion sort, comb sort when the depth budget 2*ceil(log2 n) runs out).  BWA-MEM's output
 * depends on how that sort orders equal keys (chains of equal weight, hits of equal score), so the
 * oracle and the CUDA path must both reproduce it.  Written index-based for this project.
 */
#ifndef SSQO_SORT_H
#define SSQO_SORT_H
#include <stddef.h>

#define SSQO_SORT_INIT(name, type_t, LT)                                                          \
	static inline void ssqo_isort_##name(type_t *a, long lo, long hi) /* [lo,hi) */               \
	{                                                                                             \
		long i, j;                                                                                \
		for (i = lo + 1; i < hi; ++i)                                                             \
			for (j = i; j > lo && LT(a[j], a[j - 1]); --j) { type_t t = a[j]; a[j] = a[j - 1]; a[j - 1] = t; } \
	}                                                                                             \
	static inline void ssqo_combsort_##name(type_t *a, long n)                                    \
	{                                                                                             \
		const double shrink = 1.2473309501039786540366528676643;                                  \
		int swapped;                                                                              \
		long gap = n, i;                                                                          \
		do {                                                                                      \
			if (gap > 2) { gap = (long)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; }     \
			swapped = 0;                                                                          \
			for (i = 0; i < n - gap; ++i)                                                         \
				if (LT(a[i + gap], a[i])) { type_t t = a[i]; a[i] = a[i + gap]; a[i + gap] = t; swapped = 1; } \
		} while (swapped || gap > 2);                                                             \
		if (gap != 1) ssqo_isort_##name(a, 0, n);                                                 \
	}                                                                                             \
	static inline void ssqo_introsort_##name(size_t n_, type_t *a)                                \
	{                                                                                             \
		long n = (long)n_, s, t, i, j, k, top = 0;                                                \
		int d;                                                                                    \
		struct { long l, r; int d; } stk[128];                                                    \
		if (n < 1) return;                                                                        \
		if (n == 2) { if (LT(a[1], a[0])) { type_t x = a[0]; a[0] = a[1]; a[1] = x; } return; }   \
		for (d = 2; (1L << d) < n; ++d);                                                          \
		s = 0; t = n - 1; d <<= 1;                                                                \
		for (;;) {                                                                                \
			if (s < t) {                                                                          \
				type_t rp, x;                                                                     \
				if (--d == 0) { ssqo_combsort_##name(a + s, t - s + 1); t = s; continue; }        \
				i = s; j = t; k = i + ((j - i) >> 1) + 1;                                         \
				if (LT(a[k], a[i])) { if (LT(a[k], a[j])) k = j; }                                \
				else k = LT(a[j], a[i]) ? i : j;                                                  \
				rp = a[k];                                                                        \
				if (k != t) { x = a[k]; a[k] = a[t]; a[t] = x; }                                  \
				for (;;) {                                                                        \
					do ++i; while (LT(a[i], rp));                                                 \
					do --j; while (i <= j && LT(rp, a[j]));                                       \
					if (j <= i) break;                                                            \
					x = a[i]; a[i] = a[j]; a[j] = x;                                              \
				}                                                                                 \
				x = a[i]; a[i] = a[t]; a[t] = x;                                                  \
				if (i - s > t - i) {                                                              \
					if (i - s > 16) { stk[top].l = s; stk[top].r = i - 1; stk[top].d = d; ++top; } \
					s = t - i > 16 ? i + 1 : t;                                                   \
				} else {                                                                          \
					if (t - i > 16) { stk[top].l = i + 1; stk[top].r = t; stk[top].d = d; ++top; } \
					t = i - s > 16 ? i - 1 : s;                                                   \
				}                                                                                 \
			} else {                                                                              \
				if (top == 0) { ssqo_isort_##name(a, 0, n); return; }                             \
				--top; s = stk[top].l; t = stk[top].r; d = stk[top].d;                            \
			}                                                                                     \
		}                                                                                         \
	}
#endif
