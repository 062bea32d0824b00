// ssq_pipe_host.h — the two host-side steps of the HBM-resident `bwa mem` pipeline (ssq_pipe.cu), shared with tests/hostsim:
// the reduction of the batch's insert-size histogram to mem_pestat's statistics, and the table of pairing penalties over the
// integer insert sizes those statistics admit.  Both are IEEE-double arithmetic with libm calls (sqrt, log, erfc) that must match
// the reference's evaluation bit for bit, which is why they stay on the host (speedseq:438).
#pragma once
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "ssq_dev3.cuh"

// insert-size statistics from the histogram: the reference sorts the insert sizes and sums them in that order; a histogram
// determines the sorted array, so walking it value by value replays exactly the same double operations (mem_pestat)
static void pestat_from_hist(const ssq_opts_t &o, const u32 *hist, int hist_n, PeStat pes[4], FILE *log)
{
	u64 cnt[4];
	memset(pes, 0, 4 * sizeof(PeStat));
	for (int d = 0; d < 4; ++d) { cnt[d] = 0; for (int v = 0; v < hist_n; ++v) cnt[d] += hist[(size_t)d * hist_n + v]; }
	if (log) fprintf(log, "[M::mem_pestat] # candidate unique pairs for (FF, FR, RF, RR): (%ld, %ld, %ld, %ld)\n", (long)cnt[0], (long)cnt[1], (long)cnt[2], (long)cnt[3]);
	for (int d = 0; d < 4; ++d) {
		PeStat *r = &pes[d];
		const u32 *h = hist + (size_t)d * hist_n;
		if (cnt[d] < 10) { if (log) fprintf(log, "[M::mem_pestat] skip orientation %c%c as there are not enough pairs\n", "FR"[d >> 1 & 1], "FR"[d & 1]); r->failed = 1; continue; }
		if (log) fprintf(log, "[M::mem_pestat] analyzing insert size distribution for orientation %c%c...\n", "FR"[d >> 1 & 1], "FR"[d & 1]);
		auto kth = [&](u64 k) { u64 acc = 0; for (int v = 0; v < hist_n; ++v) { acc += h[v]; if (acc > k) return v; } return hist_n - 1; };
		const int p25 = kth((u64)(int)(.25 * cnt[d] + .499)), p50 = kth((u64)(int)(.50 * cnt[d] + .499)), p75 = kth((u64)(int)(.75 * cnt[d] + .499));
		int x = 0;
		r->low = (int)(p25 - 2.0 * (p75 - p25) + .499);
		if (r->low < 1) r->low = 1;
		r->high = (int)(p75 + 2.0 * (p75 - p25) + .499);
		if (log) fprintf(log, "[M::mem_pestat] (25, 50, 75) percentile: (%d, %d, %d)\n[M::mem_pestat] low and high boundaries for computing mean and std.dev: (%d, %d)\n", p25, p50, p75, r->low, r->high);
		r->avg = 0;
		for (int v = 0; v < hist_n; ++v) if ((u64)v >= (u64)r->low && (u64)v <= (u64)r->high) for (u32 c = 0; c < h[v]; ++c) { r->avg += (u64)v; ++x; }
		r->avg /= x;
		r->std = 0;
		for (int v = 0; v < hist_n; ++v) if ((u64)v >= (u64)r->low && (u64)v <= (u64)r->high) for (u32 c = 0; c < h[v]; ++c) r->std += ((u64)v - r->avg) * ((u64)v - r->avg);
		r->std = sqrt(r->std / x);
		if (log) fprintf(log, "[M::mem_pestat] mean and std.dev: (%.2f, %.2f)\n", r->avg, r->std);
		r->low = (int)(p25 - 3.0 * (p75 - p25) + .499);
		r->high = (int)(p75 + 3.0 * (p75 - p25) + .499);
		if (r->low > r->avg - 4.0 * r->std) r->low = (int)(r->avg - 4.0 * r->std + .499);
		if (r->high < r->avg + 4.0 * r->std) r->high = (int)(r->avg + 4.0 * r->std + .499);
		if (r->low < 1) r->low = 1;
		if (log) fprintf(log, "[M::mem_pestat] low and high boundaries for proper pairs: (%d, %d)\n", r->low, r->high);
	}
	u64 max = 0;
	for (int d = 0; d < 4; ++d) max = max > cnt[d] ? max : cnt[d];
	for (int d = 0; d < 4; ++d)
		if (pes[d].failed == 0 && cnt[d] < max * 0.05) { pes[d].failed = 1; if (log) fprintf(log, "[M::mem_pestat] skip orientation %c%c\n", "FR"[d >> 1 & 1], "FR"[d & 1]); }
}


// pen[pat[d] + k] = .721 * log(2 * erfc(|dist - avg| / std / sqrt 2)) * a for dist = pes[d].low + k (upstream mem_pair's score
// adjustment, evaluated here once per admissible integer distance instead of once per candidate pair); returns the total size
static size_t pen_table(const ssq_opts_t &o, const PeStat pes[4], std::vector<double> &pen, int pn[4], size_t pat[4])
{
	size_t tot = 0;
	for (int d = 0; d < 4; ++d) { pn[d] = (!pes[d].failed && pes[d].high >= pes[d].low) ? pes[d].high - pes[d].low + 1 : 0; if (pn[d] > (1 << 22)) pn[d] = 1 << 22; pat[d] = tot; tot += pn[d]; }
	pen.assign(tot + 1, 0.);
	for (int d = 0; d < 4; ++d)
		for (int k = 0; k < pn[d]; ++k) {
			const i64 dist = (i64)pes[d].low + k;
			const double ns = (dist - pes[d].avg) / pes[d].std;
			pen[pat[d] + k] = .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * o.a;
		}
	return tot;
}
