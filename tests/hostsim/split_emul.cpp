// Host check of the half-segment form of the warp-wide striped local SW (ssq_warp.cuh: sw_local_pass_warp_split): the 16 byte lanes
// of the striped kernel own a segment of SLEN cells each; here every segment is shared by TWO warp lanes (lane s: cells
// [0, HA), lane 16 + s: cells [HA, SLEN)), so that all 32 lanes of the warp work.  The second half of a segment starts its row
// with F = 0 and repairs its cells when the first half's F arrives (a max-plus recurrence: F_true = max(F_local, F_in decayed)).
// This file runs that algorithm lane by lane in plain C++ — every shuffle an array read — against the scalar restatement of the
// striped kernel (sw_local_pass, ssq_dev2.cuh) on random problems.  build: see tests/test_split_emul.py
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../speedseq_b200/csrc/ssq_dev.cuh"
#include "../../speedseq_b200/csrc/ssq_dev2.cuh"

static unsigned long long n_rows, n_fix, n_lazy_past0, n_lazy_b, n_lazy_round2;
template <int SLEN>
static LocalRes split_pass(const ssq_opts_t &o, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int xtra, u64 *b, int b_cap)
{
	const int HA = (SLEN + 1) / 2, HB = SLEN / 2, slen = SLEN;
	const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins, e_del = o.e_del, e_ins = o.e_ins;
	const int shift = o.b > 1 ? o.b : 1, maxsc = o.a;
	const int minsc = (xtra & SSQ_XSUBO) ? xtra & 0xffff : 0x10000;
	const int endsc = (xtra & SSQ_XSTOP) ? xtra & 0xffff : 0x10000;
	LocalRes r;
	r.score = 0; r.te = r.qe = -1; r.score2 = -1; r.te2 = -1; r.tb = r.qb = -1;
	int H[32][HA], E[32][HA], HM[32][HA]; u32 PF[32][HA];
	int nloc[32], base[32];
	for (int lane = 0; lane < 32; ++lane) {
		const int s = lane & 15, half = lane >> 4;
		nloc[lane] = half ? HB : HA; base[lane] = half ? HA : 0;
		for (int j = 0; j < HA; ++j) {
			const int pos = s * slen + base[lane] + j;
			const int qc = (j < nloc[lane] && pos < qlen) ? q[pos] : -1;
			u32 w = 0;
			for (int c = 0; c < 4; ++c) w |= (u32)(uint8_t)(int8_t)(qc < 0 ? 0 : score_of(o, qc, c)) << (8 * c);
			PF[lane][j] = w; H[lane][j] = E[lane][j] = HM[lane][j] = 0;
		}
	}
	int gmax = 0, te = -1, n_b = 0;
	for (int i = 0; i < tlen; ++i) {
		const int tb = t[i], sh = 8 * tb;
		int lastH[32], hd0[32], f[32], imax[32], fin[32];
		for (int lane = 0; lane < 32; ++lane) lastH[lane] = H[lane][nloc[lane] - 1];
		for (int lane = 0; lane < 32; ++lane) { const int s = lane & 15, half = lane >> 4; hd0[lane] = lane == 0 ? 0 : lastH[half ? s : 16 + s - 1]; }
		for (int lane = 0; lane < 32; ++lane) { // local pass: exact for the first halves, provisional for the second
			int hd = hd0[lane], ff = 0, im = 0;
			for (int j = 0; j < nloc[lane]; ++j) {
				const int hn = H[lane][j];
				int h = hd + (int)(int8_t)(PF[lane][j] >> sh), e = E[lane][j], tt;
				h += shift; if (h > 255) h = 255; h -= shift; if (h < 0) h = 0;
				h = h > e ? h : e;
				h = h > ff ? h : ff;
				im = im > h ? im : h;
				H[lane][j] = h;
				tt = h - oe_del; if (tt < 0) tt = 0;
				e -= e_del; if (e < 0) e = 0;
				E[lane][j] = e > tt ? e : tt;
				tt = h - oe_ins; if (tt < 0) tt = 0;
				ff -= e_ins; if (ff < 0) ff = 0;
				ff = ff > tt ? ff : tt;
				hd = hn;
			}
			f[lane] = ff; imax[lane] = im;
		}
		bool any_fix = false;
		for (int lane = 0; lane < 32; ++lane) { fin[lane] = lane >= 16 ? f[lane & 15] : 0; if (fin[lane] > 0) any_fix = true; }
		++n_rows; if (any_fix) ++n_fix;
		if (any_fix) for (int lane = 16; lane < 32; ++lane) { // the first half's F reaches into the second half
			int g = fin[lane];
			for (int j = 0; j < nloc[lane]; ++j) {
				if (g > H[lane][j]) { H[lane][j] = g; int tt = g - oe_del; if (tt < 0) tt = 0; if (tt > E[lane][j]) E[lane][j] = tt; if (g > imax[lane]) imax[lane] = g; }
				g -= e_ins; if (g < 0) g = 0;
			}
			if (g > f[lane]) f[lane] = g;
		}
		{ // lazy F: first halves, then second halves; the exit test after each cell looks at the 16 lanes that own it
			bool done = false;
			int fl[32];
			for (int lane = 0; lane < 32; ++lane) fl[lane] = f[lane]; // only the second halves' (= the segments') F matters
			for (int round = 0; round < 16 && !done; ++round) {
				int in[32];
				for (int lane = 0; lane < 16; ++lane) in[lane] = lane == 0 ? 0 : fl[16 + lane - 1];
				for (int lane = 0; lane < 16; ++lane) fl[lane] = in[lane];
				for (int j = 0; j < HA && !done; ++j) {
					bool any = false;
					for (int lane = 0; lane < 16; ++lane) {
						int h = H[lane][j], tt;
						h = h > fl[lane] ? h : fl[lane];
						H[lane][j] = h;
						tt = h - oe_ins; if (tt < 0) tt = 0;
						fl[lane] -= e_ins; if (fl[lane] < 0) fl[lane] = 0;
						if (fl[lane] > tt) any = true;
					}
					if (!any) done = true; else if (j == 0 && round == 0) ++n_lazy_past0;
				}
				if (done) break;
				++n_lazy_b; if (round) ++n_lazy_round2;
				for (int lane = 16; lane < 32; ++lane) fl[lane] = fl[lane - 16];
				for (int j = 0; j < HB && !done; ++j) {
					bool any = false;
					for (int lane = 16; lane < 32; ++lane) {
						int h = H[lane][j], tt;
						h = h > fl[lane] ? h : fl[lane];
						H[lane][j] = h;
						tt = h - oe_ins; if (tt < 0) tt = 0;
						fl[lane] -= e_ins; if (fl[lane] < 0) fl[lane] = 0;
						if (fl[lane] > tt) any = true;
					}
					if (!any) done = true;
				}
			}
		}
		int im = 0;
		for (int lane = 0; lane < 32; ++lane) im = im > imax[lane] ? im : imax[lane];
		if (im >= minsc) {
			if (n_b == 0 || (i32)b[n_b - 1] + 1 != i) { if (n_b < b_cap) b[n_b++] = (u64)im << 32 | (u32)i; }
			else if ((int)(b[n_b - 1] >> 32) < im) b[n_b - 1] = (u64)im << 32 | (u32)i;
		}
		if (im > gmax) {
			gmax = im; te = i;
			for (int lane = 0; lane < 32; ++lane) for (int j = 0; j < HA; ++j) HM[lane][j] = H[lane][j];
			if (gmax + shift >= 255 || gmax >= endsc) break;
		}
	}
	r.score = gmax + shift < 255 ? gmax : 255;
	r.te = te;
	if (r.score != 255) {
		int vmax = -1, qe = 0x7fffffff;
		for (int lane = 0; lane < 32; ++lane) for (int j = 0; j < nloc[lane]; ++j) {
			const int v = HM[lane][j], pos = (lane & 15) * slen + base[lane] + j;
			if (v > vmax) { vmax = v; qe = pos; } else if (v == vmax && pos < qe) qe = pos;
		}
		r.qe = qe;
		if (n_b) {
			const int d = (r.score + maxsc - 1) / maxsc, low = te - d, high = te + d;
			for (int i = 0; i < n_b; ++i) { const int e = (i32)b[i]; if ((e < low || e > high) && (int)(b[i] >> 32) > r.score2) { r.score2 = (int)(b[i] >> 32); r.te2 = e; } }
		}
	}
	return r;
}

static unsigned long long rng_s = 88172645463325252ull;
static unsigned rnd() { rng_s ^= rng_s << 13; rng_s ^= rng_s >> 7; rng_s ^= rng_s << 17; return (unsigned)(rng_s >> 11); }

int main(int argc, char **argv)
{
	const int n_cases = argc > 1 ? atoi(argv[1]) : 20000;
	ssq_opts_t o; memset(&o, 0, sizeof o);
	o.a = 1; o.b = 4; o.o_del = 6; o.e_del = 1; o.o_ins = 6; o.e_ins = 1;
	int bad = 0, lazy_cases = 0;
	for (int c = 0; c < n_cases && bad < 5; ++c) {
		if (c % 7 == 3) { o.o_ins = 2 + rnd() % 5; o.e_ins = 1 + rnd() % 2; o.o_del = 2 + rnd() % 5; o.e_del = 1 + rnd() % 2; o.b = 2 + rnd() % 4; o.a = 1 + rnd() % 2; }
		else { o.a = 1; o.b = 4; o.o_del = o.o_ins = 6; o.e_del = o.e_ins = 1; }
		const int qlen = 17 + rnd() % 239, tlen = 30 + rnd() % 900;
		std::vector<uint8_t> q(qlen), t(tlen);
		const int mode = rnd() % 5;
		for (int i = 0; i < tlen; ++i) t[i] = mode == 3 ? (i / 3) % 2 : rnd() & 3; // mode 3: low complexity
		for (int i = 0; i < qlen; ++i) q[i] = mode == 3 ? (i / 3) % 2 : rnd() & 3;
		if (mode <= 2 && tlen > qlen / 2) { // plant (part of) the query with substitutions and indels: long gaps light up F and the lazy loop
			int tp = rnd() % (tlen - qlen / 2), qp = 0;
			while (qp < qlen && tp < tlen) {
				const unsigned x = rnd() % 100;
				if (x < 3) { tp += 1 + rnd() % (mode == 2 ? 12 : 3); continue; }       // deletion from the query
				if (x < 6) { qp += 1 + rnd() % (mode == 2 ? 12 : 3); continue; }       // insertion
				t[tp] = x < 10 ? rnd() & 3 : q[qp]; ++tp; ++qp;
			}
		}
		if (rnd() % 10 == 0) q[rnd() % qlen] = 4; // an N in the query (profile 'N' column: -1)
		const int xt = rnd() % 3;
		const int xtra = SSQ_XBYTE | (xt == 1 ? SSQ_XSUBO | (rnd() % 30 + 10) : xt == 2 ? SSQ_XSTOP | (rnd() % 60 + 20) : 0);
		const int slen = (qlen + 15) / 16, n = slen * 16;
		std::vector<i32> h0(n), h1(n), e(n), hm(n);
		std::vector<u64> b1(tlen + 1), b2(tlen + 1);
		LocalScratch S; S.H0 = h0.data(); S.H1 = h1.data(); S.E = e.data(); S.Hmax = hm.data(); S.b = b1.data(); S.b_cap = tlen + 1;
		const LocalRes want = sw_local_pass(o, true, qlen, q.data(), tlen, t.data(), xtra, S);
		LocalRes got;
		switch (slen) {
#define C(n_) case n_: got = split_pass<n_>(o, qlen, q.data(), tlen, t.data(), xtra, b2.data(), tlen + 1); break;
			C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16)
#undef C
			default: continue;
		}
		if (want.score > 60) ++lazy_cases;
		if (memcmp(&want, &got, sizeof want) != 0) {
			++bad;
			fprintf(stderr, "case %d (qlen %d tlen %d mode %d xtra %x): want score %d te %d qe %d score2 %d te2 %d, got %d %d %d %d %d\n", c, qlen, tlen, mode, xtra, want.score, want.te, want.qe, want.score2, want.te2,
			        got.score, got.te, got.qe, got.score2, got.te2);
		}
	}
	printf("%d cases, %d with score > 60, %d mismatches; %llu rows: %llu with a repair of the second halves, %llu where the lazy loop went past its first cell, %llu sweeps into the second halves, %llu of them in later rounds\n", n_cases, lazy_cases, bad, n_rows, n_fix, n_lazy_past0, n_lazy_b, n_lazy_round2);
	return bad ? 1 : 0;
}
