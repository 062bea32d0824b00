// hostsim.cpp — TEST-ONLY harness: compiles the SSQ_HD routines of speedseq_b200/csrc/ssq_dev.cuh for the host and runs
// them per read exactly as the kernels' thread/warp bodies do (ScalarFm stands in for the warp-cooperative rank query).
// It exists so that `pytest -m "not gpu"` can check the kernels' arithmetic against the oracle on a box with no GPU.
// It is NOT part of libssq.so and nothing in the product loads it; the product has no CPU path.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../speedseq_b200/csrc/ssq_dev.cuh"
extern "C" {
#include "../../oracle/ssqo.h"
}

static DevIndex make_ix(const ssqo_idx_t *o, std::vector<i64> &off, std::vector<i32> &len)
{
	DevIndex ix;
	memset(&ix, 0, sizeof ix);
	ix.bwt = o->bwt.bwt; ix.sa = o->bwt.sa; ix.pac = o->pac;
	ix.primary = o->bwt.primary; memcpy(ix.L2, o->bwt.L2, sizeof ix.L2); ix.seq_len = o->bwt.seq_len; ix.n_sa = o->bwt.n_sa;
	ix.l_pac = o->bns.l_pac; ix.n_seqs = o->bns.n_seqs; ix.sa_intv = o->bwt.sa_intv;
	off.resize(ix.n_seqs); len.resize(ix.n_seqs);
	for (int i = 0; i < ix.n_seqs; ++i) { off[i] = o->bns.anns[i].offset; len[i] = o->bns.anns[i].len; }
	ix.ann_off = off.data(); ix.ann_len = len.data();
	return ix;
}

struct ReadWork {
	std::vector<Intv> mem; int n_intv, l_rep;
	std::vector<Seed> seeds, sorted; std::vector<ChainRec> outc; int n_kept;
	std::vector<RegCand> regs;
};

static void run_read(const DevIndex &ix, const ssq_opts_t &opt, int len, const uint8_t *q, int upto, ReadWork &w)
{
	ScalarFm fm(ix);
	std::vector<Intv> bufA(len + 2), bufB(len + 2);
	w.mem.assign(2048, Intv());
	int err = 0;
	if (getenv("HOSTSIM_STRAIGHT")) w.n_intv = collect_intv(fm, ix, opt, len, q, w.mem.data(), 2048, bufA.data(), bufB.data(), err);
	else { // the state-machine form the GPU kernel runs
		SmemMachine m; Intv ok[4];
		m.init(opt, len, q, w.mem.data(), 2048, bufA.data(), bufB.data());
		while (m.advance(ix)) { fm.extend(m.in, ok, m.is_back); m.post(ok); }
		err = m.err; w.n_intv = m.finish();
	}
	if (err) abort();
	int b = 0, en = 0; w.l_rep = 0;
	for (int i = 0; i < w.n_intv; ++i) {
		const Intv p = w.mem[i];
		if (p.x2 <= (u64)opt.max_occ) continue;
		if ((int)p.qb > en) { w.l_rep += en - b; b = p.qb; en = p.qe; } else en = en > (int)p.qe ? en : (int)p.qe;
	}
	w.l_rep += en - b;
	w.n_kept = 0; w.seeds.clear(); w.regs.clear();
	if (upto < 1) return;
	unsigned long long n_sa = 0;
	for (int i = 0; i < w.n_intv; ++i) {
		u64 step; int cnt = intv_occ_count(w.mem[i].x2, opt.max_occ, step);
		for (int k = 0; k < cnt; ++k) {
			Seed s; s.rbeg = (i64)sa_lookup(fm, w.mem[i].x0 + (u64)k * step, n_sa); s.qbeg = w.mem[i].qb; s.len = w.mem[i].qe - w.mem[i].qb;
			w.seeds.push_back(s);
		}
	}
	int n = (int)w.seeds.size();
	if (n == 0) return;
	std::vector<i32> chain_of(n), ord(n); std::vector<ChainRec> ch(n); std::vector<WIdx> wi(n);
	w.sorted.assign(n, Seed()); w.outc.assign(n, ChainRec());
	w.n_kept = chain_and_filter(ix, opt, len, n, w.seeds.data(), w.l_rep, chain_of.data(), ch.data(), ord.data(), wi.data(), w.sorted.data(), w.outc.data());
	if (upto < 2) return;
	std::vector<u32> ehbuf(len + 4);
	EhAcc eh; eh.base = ehbuf.data(); eh.stride = 1;
	std::vector<RegCand> out; int n_out = 0;
	size_t total = 0;
	for (int c = 0; c < w.n_kept; ++c) total += w.outc[c].n;
	out.resize(total + 1);
	for (int c = 0; c < w.n_kept; ++c) {
		const ChainRec &cr = w.outc[c];
		std::vector<RegCand> cand(cr.n); std::vector<u64> srt(cr.n);
		for (int s = 0; s < cr.n; ++s) extend_seed(ix, opt, len, q, cr, w.sorted.data() + cr.seed_start, s, eh, cand[s], 0);
		select_regions(opt, len, cr, w.sorted.data() + cr.seed_start, cand.data(), srt.data(), out.data(), n_out);
	}
	w.regs.assign(out.begin(), out.begin() + n_out);
}

extern "C" {

typedef struct { uint64_t k, l, s; uint32_t qbeg, qend; } api_smem_t;
typedef struct { int64_t rbeg; int32_t qbeg, len; } api_seed_t;
typedef struct { int64_t rb, re; int32_t qb, qe, rid, score, truesc, w, seedcov, seedlen0; float frac_rep; int32_t read_id; } api_alnreg_t;
typedef struct { uint64_t q_off, t_off; int32_t qlen, tlen, h0, w, end_bonus, zdrop; } api_sw_task_t;
typedef struct { int32_t score, qle, tle, gtle, gscore, max_off; } api_sw_result_t;

int64_t hostsim_smem_batch(const ssqo_idx_t *idx, int n_reads, const uint8_t *seq, const uint64_t *read_off, api_smem_t *out, uint64_t cap, uint64_t *out_off)
{
	std::vector<i64> off; std::vector<i32> len; DevIndex ix = make_ix(idx, off, len);
	ssq_opts_t opt; ssq_opts_default(&opt);
	uint64_t n = 0;
	for (int r = 0; r < n_reads; ++r) {
		ReadWork w; out_off[r] = n;
		run_read(ix, opt, (int)(read_off[r + 1] - read_off[r]), seq + read_off[r], 0, w);
		for (int i = 0; i < w.n_intv; ++i, ++n) { if (n >= cap) return -1; out[n].k = w.mem[i].x0; out[n].l = w.mem[i].x1; out[n].s = w.mem[i].x2; out[n].qbeg = w.mem[i].qb; out[n].qend = w.mem[i].qe; }
	}
	out_off[n_reads] = n;
	return (int64_t)n;
}

int64_t hostsim_chain_batch(const ssqo_idx_t *idx, int n_reads, const uint8_t *seq, const uint64_t *read_off, api_seed_t *seeds, uint64_t seed_cap,
                            uint64_t *chain_seed_off, uint64_t chain_cap, uint64_t *read_chain_off)
{
	std::vector<i64> off; std::vector<i32> len; DevIndex ix = make_ix(idx, off, len);
	ssq_opts_t opt; ssq_opts_default(&opt);
	uint64_t ns = 0, nc = 0;
	for (int r = 0; r < n_reads; ++r) {
		ReadWork w; read_chain_off[r] = nc;
		run_read(ix, opt, (int)(read_off[r + 1] - read_off[r]), seq + read_off[r], 1, w);
		for (int c = 0; c < w.n_kept; ++c) {
			if (nc >= chain_cap) return -1;
			chain_seed_off[nc++] = ns;
			for (int s = 0; s < w.outc[c].n; ++s, ++ns) {
				if (ns >= seed_cap) return -1;
				const Seed &x = w.sorted[w.outc[c].seed_start + s];
				seeds[ns].rbeg = x.rbeg; seeds[ns].qbeg = x.qbeg; seeds[ns].len = x.len;
			}
		}
	}
	chain_seed_off[nc] = ns; read_chain_off[n_reads] = nc;
	return (int64_t)nc;
}

int64_t hostsim_align_batch(const ssqo_idx_t *idx, int n_reads, const uint8_t *seq, const uint64_t *read_off, api_alnreg_t *out, uint64_t cap, uint64_t *out_off)
{
	std::vector<i64> off; std::vector<i32> len; DevIndex ix = make_ix(idx, off, len);
	ssq_opts_t opt; ssq_opts_default(&opt);
	uint64_t n = 0;
	for (int r = 0; r < n_reads; ++r) {
		ReadWork w; out_off[r] = n;
		run_read(ix, opt, (int)(read_off[r + 1] - read_off[r]), seq + read_off[r], 2, w);
		for (size_t i = 0; i < w.regs.size(); ++i, ++n) {
			if (n >= cap) return -1;
			const RegCand &a = w.regs[i];
			out[n].rb = a.rb; out[n].re = a.re; out[n].qb = a.qb; out[n].qe = a.qe; out[n].rid = a.rid; out[n].score = a.score; out[n].truesc = a.truesc;
			out[n].w = a.w; out[n].seedcov = a.seedcov; out[n].seedlen0 = a.seedlen0; out[n].frac_rep = a.frac_rep; out[n].read_id = r;
		}
	}
	out_off[n_reads] = n;
	return (int64_t)n;
}

int hostsim_sw_extend_batch(uint64_t n, const api_sw_task_t *t, const uint8_t *qbuf, const uint8_t *tbuf, api_sw_result_t *r)
{
	ssq_opts_t opt; ssq_opts_default(&opt);
	for (uint64_t i = 0; i < n; ++i) {
		std::vector<u32> ehbuf(t[i].qlen + 4);
		EhAcc eh; eh.base = ehbuf.data(); eh.stride = 1;
		const uint8_t *q = qbuf + t[i].q_off, *tg = tbuf + t[i].t_off;
		unsigned long long cells = 0;
		r[i].score = sw_extend(opt, t[i].qlen, [&](int j) { return (int)q[j]; }, t[i].tlen, [&](int k) { return (int)tg[k]; }, t[i].w, t[i].end_bonus, t[i].zdrop, t[i].h0, eh,
		                       r[i].qle, r[i].tle, r[i].gtle, r[i].gscore, r[i].max_off, cells);
	}
	return 0;
}

void ssq_opts_default(ssq_opts_t *o) // same values as speedseq_b200/csrc/ssq_index.cu (kept in sync by tests/test_abi.py)
{
	memset(o, 0, sizeof *o);
	o->a = 1; o->b = 4; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
	o->pen_unpaired = 17; o->pen_clip5 = o->pen_clip3 = 5; o->w = 100; o->zdrop = 100; o->T = 30;
	o->min_seed_len = 19; o->split_width = 10; o->max_occ = 500; o->max_chain_gap = 10000; o->max_mem_intv = 20;
	o->min_chain_weight = 0; o->max_chain_extend = 1 << 30; o->max_ins = 10000; o->max_matesw = 50; o->max_XA_hits = 5;
	o->split_factor = 1.5f; o->mask_level = 0.50f; o->drop_ratio = 0.50f; o->XA_drop_ratio = 0.80f; o->mask_level_redun = 0.95f;
	o->mapQ_coef_len = 50; o->mapQ_coef_fac = 3;
}
}
