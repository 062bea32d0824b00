// hostsim.cpp — TEST-ONLY harness: compiles the SSQ_HD routines of speedseq_b200/csrc/ssq_dev.cuh for the host and runs
// them per read exactly as the kernels' thread/warp bodies do (ScalarFm stands in for the warp-cooperative rank query).
// It exists so that `pytest -m "not gpu"` can check the kernels' arithmetic against the oracle on a box with no GPU.
// It is NOT part of libssq.so and nothing in the product loads it; the product has no CPU path.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../../speedseq_b200/csrc/ssq_dev.cuh"
#include "../../speedseq_b200/csrc/ssq_dev3.cuh"
#include "../../speedseq_b200/csrc/ssq_pipe_host.h"
#include <set>
#include <string>
extern "C" {
#include "../../oracle/ssqo.h"
}

static std::vector<u32> g_bwt32; // re-blocked copy, rebuilt per call (test-only)
static DevIndex make_ix(const ssqo_idx_t *o, std::vector<i64> &off, std::vector<i32> &len)
{
	DevIndex ix;
	memset(&ix, 0, sizeof ix);
	if (!getenv("HOSTSIM_NO_BWT32")) { // same derivation as k_reblock32 in ssq_index.cu
		const u64 n = o->bwt.seq_len, nb = (n + 63) / 64 + 1;
		g_bwt32.assign(nb * 8, 0);
		for (u64 b = 0; b < nb; ++b) {
			const u32 *src = o->bwt.bwt + ((b >> 1) << 4);
			const u64 *c64 = (const u64*)src;
			u32 add[4] = {0, 0, 0, 0};
			if (b & 1) for (int i = 0; i < 64; ++i) ++add[src[8 + (i >> 4)] >> ((~i & 15) << 1) & 3];
			const bool have = ((b >> 1) << 7) < n + 128; // within the stored blocks (incl. the final count block)
			for (int c = 0; c < 4; ++c) g_bwt32[b * 8 + c] = have ? (u32)(c64[c] + add[c]) : 0;
			for (int w = 0; w < 4; ++w) g_bwt32[b * 8 + 4 + w] = (((b >> 1) << 7) + (b & 1) * 64 + w * 16 < n) ? src[8 + (b & 1) * 4 + w] : 0;
		}
		ix.bwt32 = g_bwt32.data();
	}
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

// The phase-split formulation of the seeding (k_smem_fwd / k_smem_bwd / k_smem_p3 on the GPU), data flow included: forward walks
// produce calls + forward lists, every call's backward phase runs on its own from that list, pass 2 is selected from pass 1's
// intervals, everything is pooled and ordered by (qb, qe) at the end.
// diagnostics (HOSTSIM_SPLIT): rank queries by the length of the string they produce — sizes a k-mer jump-start table
static unsigned long long g_ext_hist[3][256], g_tab_hits;
extern "C" void hostsim_ext_hist(unsigned long long *out) { memcpy(out, g_ext_hist, sizeof g_ext_hist); }
extern "C" void hostsim_ext_hist_reset() { memset(g_ext_hist, 0, sizeof g_ext_hist); g_tab_hits = 0; }
extern "C" unsigned long long hostsim_table_hits() { return g_tab_hits; }

// k-mer jump-start table for the diagnostics / HOSTSIM_KMER=K runs: built with the same kmer_compute() the GPU loader uses
static std::vector<KmerEnt> g_kmer; static int g_kmer_k = 0; static const void *g_kmer_for = 0;
static void ensure_kmer(DevIndex &ix)
{
	const int K = getenv("HOSTSIM_KMER") ? atoi(getenv("HOSTSIM_KMER")) : 0;
	if (K <= 0 || !ix.bwt32) { ix.kmer = 0; ix.kmer_k = 0; return; }
	if (g_kmer_k != K || g_kmer_for != (const void*)ix.bwt32) {
		g_kmer.assign(kmer_entries(K), KmerEnt());
		ScalarFm fm(ix);
		for (int L = 1; L <= K; ++L) for (u32 code = 0; code < (1u << (2 * L)); ++code) g_kmer[kmer_off(L) + code] = kmer_compute(fm, ix, L, code);
		g_kmer_k = K; g_kmer_for = (const void*)ix.bwt32;
	}
	ix.kmer = g_kmer.data(); ix.kmer_k = K;
}

static void fwd_walk(ScalarFm &fm, const DevIndex &ix, int len, const uint8_t *q, int x, u32 min_intv, std::vector<FwdEntry> &list)
{
	u32 code = q[x]; // the walk's string so far, for the k-mer table
	Intv32 ik, okc;
	set_intv(ix, q[x], ik); ik.qe = (u32)(x + 1);
	auto push = [&]() { FwdEntry e; e.x0 = ik.x0; e.x1 = ik.x1; e.x2 = ik.x2; e.qe = ik.qe; list.push_back(e); };
	int i;
	for (i = x + 1; i < len; ++i) {
		if (q[i] > 3) break;
		code = code << 2 | q[i];
		if (ix.kmer_k && i + 1 - x <= ix.kmer_k) { okc = kmer_lookup(ix, i + 1 - x, code); ++g_tab_hits; } else extend1(fm, ik, 3 - q[i], 0, okc);
		++g_ext_hist[0][i + 1 - x < 255 ? i + 1 - x : 255];
		if (okc.x2 != ik.x2) { push(); if (okc.x2 < min_intv) return; }
		ik = okc; ik.qe = (u32)(i + 1);
	}
	push();
}
static void run_read_split(const DevIndex &ix, const ssq_opts_t &opt, int len, const uint8_t *q, ReadWork &w)
{
	ScalarFm fm(ix);
	std::vector<Intv> all;
	std::vector<Intv> mem(2048);
	std::vector<Intv32> a32(len + 2), b32(len + 2);
	HostListsT<u32> hl; hl.a[0] = a32.data(); hl.a[1] = b32.data();
	const int split_len = (int)(opt.min_seed_len * opt.split_factor + .499f);
	const bool lean = getenv("HOSTSIM_SPLIT_LEAN") != 0; // BwdCallT (k_smem_bwd2) instead of the machine started in its backward phase
	auto backward = [&](int x, u32 min_intv, const std::vector<FwdEntry> &list) {
		if (lean) {
			BwdCallT<HostListsT<u32> > m; Intv32 okc;
			m.start(opt, len, q, mem.data(), 2048, hl, x, min_intv, list.data(), (int)list.size(), x >= 1 ? (int)q[x - 1] : 4, x >= 2 ? (int)q[x - 2] : 4);
			if (ix.kmer_k) m.use_table(ix.kmer_k);
			while (m.advance()) { if (!m.table_hit(ix, okc)) extend1(fm, m.in, m.c, 1, okc); else ++g_tab_hits; { const int l = (int)m.in.qe - m.i; ++g_ext_hist[1][l < 255 ? l : 255]; } m.post(okc); }
			if (m.err) abort();
			for (int k = 0; k < m.n; ++k) all.push_back(mem[k]);
			return;
		}
		SmemMachineT<HostListsT<u32>, u32, false> m; Intv32 okc;
		m.init(opt, len, q, mem.data(), 2048, hl, 1);
		m.start_backward(x, min_intv, list.data(), (int)list.size(), x >= 1 ? (int)q[x - 1] : 4, x >= 2 ? (int)q[x - 2] : 4);
		for (bool go = m.advance(ix); go; go = m.advance(ix)) { extend1(fm, m.in, m.qc, m.is_back, okc); m.post(okc); }
		if (m.err) abort();
		for (int k = 0; k < m.n; ++k) all.push_back(mem[k]);
	};
	if (len >= opt.min_seed_len) {
		for (int x = 0; x < len;) { // pass 1
			if (q[x] > 3) { ++x; continue; }
			std::vector<FwdEntry> list;
			fwd_walk(fm, ix, len, q, x, 1, list);
			backward(x, 1, list);
			x = (int)list.back().qe;
		}
		const size_t n1 = all.size();
		for (size_t k = 0; k < n1; ++k) { // pass 2
			const Intv p = all[k];
			const int start = (int)p.qb, end = (int)p.qe;
			if (end - start < split_len || p.x2 > (u64)opt.split_width) continue;
			std::vector<FwdEntry> list;
			const int x = (start + end) >> 1;
			fwd_walk(fm, ix, len, q, x, (u32)(p.x2 + 1), list);
			backward(x, (u32)(p.x2 + 1), list);
		}
		if (opt.max_mem_intv > 0) { // pass 3, as in k_smem_p3
			int x = 0;
			while (x < len) {
				if (q[x] > 3) { ++x; continue; }
				Intv32 ik, okc; set_intv(ix, q[x], ik);
				int i; bool hit = false;
				u32 code = q[x];
				for (i = x + 1; i < len; ++i) {
					if (q[i] > 3) break;
					code = code << 2 | q[i];
					if (ix.kmer_k && i + 1 - x <= ix.kmer_k) { okc = kmer_lookup(ix, i + 1 - x, code); ++g_tab_hits; } else extend1(fm, ik, 3 - q[i], 0, okc);
					++g_ext_hist[2][i + 1 - x < 255 ? i + 1 - x : 255];
					if (okc.x2 < (u32)opt.max_mem_intv && i - x >= opt.min_seed_len) { if (okc.x2 > 0) { Intv m = widen(okc); m.qb = (u32)x; m.qe = (u32)(i + 1); all.push_back(m); } hit = true; break; }
					ik = okc;
				}
				x = (hit || i < len) ? i + 1 : len;
			}
		}
	}
	std::stable_sort(all.begin(), all.end(), [](const Intv &a, const Intv &b) { return ((u64)a.qb << 32 | a.qe) < ((u64)b.qb << 32 | b.qe); });
	w.mem.assign(all.begin(), all.end()); w.mem.resize(all.size() + 2048);
	w.n_intv = (int)all.size();
}

static void run_read(const DevIndex &ix, const ssq_opts_t &opt, int len, const uint8_t *q, int upto, ReadWork &w)
{
	ScalarFm fm(ix);
	std::vector<Intv> bufA(len + 2), bufB(len + 2);
	w.mem.assign(2048, Intv());
	int err = 0;
	if ((getenv("HOSTSIM_SPLIT") || getenv("HOSTSIM_SPLIT_LEAN")) && ix.bwt32) { DevIndex ixk = ix; ensure_kmer(ixk); run_read_split(ixk, opt, len, q, w); }
	else if (getenv("HOSTSIM_STRAIGHT")) w.n_intv = collect_intv(fm, ix, opt, len, q, w.mem.data(), 2048, bufA.data(), bufB.data(), err);
	else if (ix.bwt32 && !getenv("HOSTSIM_M64") && getenv("HOSTSIM_KMER")) { // the 32-bit machine with the k-mer jump-start table (k_smem_m<u32, 6, true>)
		DevIndex ixk = ix; ensure_kmer(ixk);
		std::vector<Intv32> a32(len + 2), b32(len + 2);
		SmemMachineT<HostListsT<u32>, u32, true, true> m; Intv32 okc;
		HostListsT<u32> hl; hl.a[0] = a32.data(); hl.a[1] = b32.data();
		m.init(opt, len, q, w.mem.data(), 2048, hl);
		for (bool go = m.advance(ixk); go; go = m.advance(ixk)) { if (!m.table_hit(ixk, okc)) extend1(fm, m.in, m.qc, m.is_back, okc); else ++g_tab_hits; m.post(okc); }
		std::vector<u32> keys(2048);
		err = m.err; w.n_intv = m.finish(keys.data());
		{ std::vector<Intv> tmp(w.n_intv); for (int i = 0; i < w.n_intv; ++i) tmp[i] = w.mem[keys[i] & 0xffff]; for (int i = 0; i < w.n_intv; ++i) w.mem[i] = tmp[i]; }
	}
	else if (ix.bwt32 && !getenv("HOSTSIM_M64")) { // the state-machine form the GPU kernel runs, 32-bit rows (what the GPU picks when bwt32 exists)
		std::vector<Intv32> a32(len + 2), b32(len + 2);
		SmemMachineT<HostListsT<u32>, u32> m; Intv32 okc;
		HostListsT<u32> hl; hl.a[0] = a32.data(); hl.a[1] = b32.data();
		m.init(opt, len, q, w.mem.data(), 2048, hl);
		for (bool go = m.advance(ix); go; go = m.advance(ix)) { extend1(fm, m.in, m.qc, m.is_back, okc); m.post(okc); }
		std::vector<u32> keys(2048);
		err = m.err; w.n_intv = m.finish(keys.data());
		{ std::vector<Intv> tmp(w.n_intv); for (int i = 0; i < w.n_intv; ++i) tmp[i] = w.mem[keys[i] & 0xffff]; for (int i = 0; i < w.n_intv; ++i) w.mem[i] = tmp[i]; }
	} else { // 64-bit rows
		SmemMachineT<HostLists> m; Intv okc;
		HostLists hl; hl.a[0] = bufA.data(); hl.a[1] = bufB.data();
		m.init(opt, len, q, w.mem.data(), 2048, hl);
		for (bool go = m.advance(ix); go; go = m.advance(ix)) { extend1(fm, m.in, m.qc, m.is_back, okc); m.post(okc); }
		std::vector<u32> keys(2048);
		err = m.err; w.n_intv = m.finish(keys.data());
		{ std::vector<Intv> tmp(w.n_intv); for (int i = 0; i < w.n_intv; ++i) tmp[i] = w.mem[keys[i] & 0xffff]; for (int i = 0; i < w.n_intv; ++i) w.mem[i] = tmp[i]; }
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
	std::vector<i32> chain_of(n), ord(n); std::vector<ChainRec> ch(n); std::vector<WIdx> wi(n); std::vector<KeptChain> kp(n);
	w.sorted.assign(n, Seed()); w.outc.assign(n, ChainRec());
	w.n_kept = chain_and_filter(ix, opt, len, n, w.seeds.data(), w.l_rep, chain_of.data(), ch.data(), ord.data(), wi.data(), w.sorted.data(), w.outc.data(), kp.data());
	if (upto < 2) return;
	std::vector<u32> ehbuf(len + 4);
	EhAcc eh; eh.base = ehbuf.data(); eh.stride = 1;
	std::vector<RegCand> out; int n_out = 0;
	size_t total = 0;
	for (int c = 0; c < w.n_kept; ++c) total += w.outc[c].n;
	out.resize(total + 1);
	if (getenv("HOSTSIM_EAGER")) { // extend every seed, then replay
		for (int c = 0; c < w.n_kept; ++c) {
			const ChainRec &cr = w.outc[c];
			std::vector<RegCand> cand(cr.n); std::vector<u64> srt(cr.n);
			for (int s = 0; s < cr.n; ++s) extend_seed(ix, opt, len, q, cr, w.sorted.data() + cr.seed_start, s, eh, cand[s], 0);
			select_regions(opt, len, cr, w.sorted.data() + cr.seed_start, cand.data(), srt.data(), out.data(), n_out);
		}
	} else { // the GPU's lazy rounds: longest seed of every chain first, then only what the replay asks for
		std::vector<RegCand> cand(total + 1); std::vector<uint8_t> have(total + 1, 0), need(total + 1, 0); std::vector<u64> srt(total + 1);
		size_t t = 0;
		for (int c = 0; c < w.n_kept; ++c) {
			const ChainRec &cr = w.outc[c]; const Seed *cs = w.sorted.data() + cr.seed_start;
			int best = 0;
			for (int i = 1; i < cr.n; ++i) if (cs[i].len >= cs[best].len) best = i;
			need[t + best] = 1; t += cr.n;
		}
		SelState st; memset(&st, 0, sizeof(st));
		for (int round = 0;; ++round) {
			t = 0;
			for (int c = 0; c < w.n_kept; ++c) {
				const ChainRec &cr = w.outc[c];
				for (int s = 0; s < cr.n; ++s) if (need[t + s] && !have[t + s]) { extend_seed(ix, opt, len, q, cr, w.sorted.data() + cr.seed_start, s, eh, cand[t + s], 0); have[t + s] = 1; }
				t += cr.n;
			}
			// the same resumable per-read step the GPU's k_select takes each round
			const bool complete = select_read(opt, len, w.n_kept, w.outc.data(), w.sorted.data(), cand.data(), srt.data(), out.data(), have.data(), need.data(), round >= 2, st);
			n_out = st.n_out;
			if (complete) break;
		}
	}
	w.regs.assign(out.begin(), out.begin() + n_out);
}

// ---- the HBM-resident `bwa mem | samblaster` pipeline of ssq_pipe.cu, stage by stage, with the kernels' per-thread bodies
// (ssq_dev3.cuh) run in plain loops over host vectors.  Same order of stages, same host-side reductions (ssq_pipe_host.h). ----
struct HostPipe {
	std::set<std::pair<u64, u64> > seen; // the streaming dup-set: signatures of every earlier batch
	std::string text[3], bam[3]; // bam: coordinate-sorted BAM records of the batch (filled when want_bam)
	int want_bam = 0, bam_blank_side = 1;
	std::vector<u64> read_off;
	PeStat pes[4];
	int err;
	int run(const DevIndex &ix, const ssq_opts_t &opt, const SbOpts &sb, int n, const char *const *names, const char *const *seqs, const char *const *quals, const char *const *comments,
	        i64 n_processed, int paired, const PeStat *pes0, const char *rg_id, const std::vector<std::string> &ctg, const std::vector<i32> &ctg_len)
	{
		err = 0;
		// inputs as the aligner lays them out
		std::vector<u64> off(n + 1, 0); std::vector<u32> name_off(n + 1, 0), cmt_off(n + 1, 0);
		std::string nameb, qualb, cmtb;
		for (int i = 0; i < n; ++i) { off[i + 1] = off[i] + strlen(seqs[i]); nameb += names[i]; name_off[i + 1] = (u32)nameb.size(); if (quals && quals[i]) qualb += quals[i]; if (comments && comments[i]) cmtb += comments[i]; cmt_off[i + 1] = (u32)cmtb.size(); }
		std::vector<uint8_t> codes(off[n] + 16);
		for (int i = 0; i < n; ++i) for (size_t k = 0; seqs[i][k]; ++k) { const int c = seqs[i][k] | 0x20; codes[off[i] + k] = c == 'a' ? 0 : c == 'c' ? 1 : c == 'g' ? 2 : c == 't' ? 3 : 4; }
		std::string ctgb; std::vector<u32> ctg_off(ctg.size() + 1, 0); std::vector<i64> sb_off(ctg.size() + 1, 0);
		{ i64 total = 0; for (size_t i = 0; i < ctg.size(); ++i) { ctgb += ctg[i]; ctg_off[i + 1] = (u32)ctgb.size(); sb_off[i] = total; total += (i64)ctg_len[i] + 2 * SB_PAD + 1; } }
		std::vector<double> logn(8192); std::vector<i32> lg(65536);
		for (int i = 0; i < 8192; ++i) logn[i] = i ? log((double)i) : 0.;
		for (int i = 0; i < 65536; ++i) lg[i] = (int)(4.343 * log((double)(i + 1)) + .499);
		// stage 0: regions out of seed extension
		std::vector<RegCand> regs; std::vector<u64> task_off(n + 1, 0); std::vector<u32> n_regs(n + 1, 0);
		for (int r = 0; r < n; ++r) { ReadWork w; run_read(ix, opt, (int)(off[r + 1] - off[r]), codes.data() + off[r], 2, w); task_off[r] = regs.size(); n_regs[r] = (u32)w.regs.size(); regs.insert(regs.end(), w.regs.begin(), w.regs.end()); }
		regs.push_back(RegCand());
		PipeView V; memset(&V, 0, sizeof V);
		V.ix = ix; V.opt = opt; V.sb = sb;
		V.T.logn = logn.data(); V.T.n_logn = 8192; V.T.lg4343 = lg.data(); V.T.n_lg = 65536;
		V.tc.ctg_names = ctgb.data(); V.tc.ctg_name_off = ctg_off.data(); V.tc.names = nameb.data(); V.tc.name_off = name_off.data();
		V.tc.seq = codes.data(); V.tc.read_off = off.data(); V.tc.qual = quals && qualb.size() == off[n] && n ? qualb.data() : 0;
		V.tc.cmt = comments ? cmtb.data() : 0; V.tc.cmt_off = comments ? cmt_off.data() : 0;
		V.tc.rg_id = rg_id ? rg_id : ""; V.tc.rg_len = rg_id ? (i32)strlen(rg_id) : 0;
		V.n_reads = n; V.paired = paired; V.n_processed = n_processed;
		V.task_off = task_off.data(); V.n_regs = n_regs.data(); V.regs = regs.data();
		V.sb_off = sb_off.data(); V.err = &err;
		const int n_pairs = paired ? n / 2 : 0, n_units = paired ? n / 2 : n;
		// region lists
		std::vector<u64> aoff(n + 1, 0); std::vector<u32> na(n + 1, 0);
		for (int i = 0; i < n; ++i) { u32 m = paired ? n_regs[i ^ 1] : 0; if (m > (u32)opt.max_matesw) m = (u32)opt.max_matesw; aoff[i + 1] = aoff[i] + n_regs[i] + 4ull * m + (paired ? 4 : 0); }
		std::vector<AlnReg> areg(aoff[n] + 1); std::vector<P64> pv(aoff[n] + 2); std::vector<i32> xcnt(aoff[n] + 2);
		V.areg_off = aoff.data(); V.areg = areg.data(); V.n_areg = na.data(); V.pv = pv.data(); V.xcnt = xcnt.data();
		std::vector<uint8_t> qbuf(256), rbuf(2048), zb(256 * 768), sq(256 + 64); std::vector<i32> h(272), e(272), H0(272), H1(272), E(272), Hm(272);
		AlnScratch A; A.qbuf = qbuf.data(); A.rbuf = rbuf.data(); A.rcap = 2048; A.g.h = h.data(); A.g.e = e.data(); A.g.z = 0; A.g.zcap = 0;
		for (int r = 0; r < n; ++r) body_dedup(V, r, A);
		// insert-size statistics
		memset(pes, 0, sizeof pes);
		V.pes = pes;
		std::vector<double> pen; int pn[4]; size_t pat[4];
		if (paired) {
			if (pes0) for (int d = 0; d < 4; ++d) pes[d] = pes0[d];
			else {
				const int hist_n = opt.max_ins + 1;
				std::vector<u32> hist((size_t)4 * hist_n, 0);
				V.hist = hist.data(); V.hist_n = hist_n;
				for (int p = 0; p < n_pairs; ++p) { int dir; i64 is; if (body_pestat(V, p, &dir, &is)) ++hist[(size_t)dir * hist_n + is]; }
				pestat_from_hist(opt, hist.data(), hist_n, pes, 0);
			}
			pen_table(opt, pes, pen, pn, pat);
			for (int d = 0; d < 4; ++d) { V.T.pen[d] = pen.data() + pat[d]; V.T.pen_low[d] = pes[d].low; V.T.pen_n[d] = pn[d]; }
			// mate rescue
			int win = 0;
			for (int d = 0; d < 4; ++d) if (!pes[d].failed && pes[d].high - pes[d].low > win) win = pes[d].high - pes[d].low;
			int max_len = 0; for (int i = 0; i < n; ++i) if ((int)(off[i + 1] - off[i]) > max_len) max_len = (int)(off[i + 1] - off[i]);
			const int win_cap = win + max_len + 16;
			std::vector<uint8_t> ref(win_cap); std::vector<u64> bl(win_cap); std::vector<AlnReg> bbuf(128);
			MateScratch M; M.seq = sq.data(); M.ref = ref.data(); M.ref_cap = win_cap; M.L.H0 = H0.data(); M.L.H1 = H1.data(); M.L.E = E.data(); M.L.Hmax = Hm.data(); M.L.b = bl.data(); M.L.b_cap = win_cap;
			M.A.qbuf = M.A.rbuf = 0; M.A.rcap = 0; M.A.g.h = M.A.g.e = 0; M.A.g.z = 0; M.A.g.zcap = 0;
			// the device's three phases (ssq_pipe.cu): count and list the rescue alignments the initial lists do not skip, compute them
			// all ahead, then replay every pair in the reference's order with the results looked up.  HOSTSIM_RESCUE_SPEC=0: the plain replay
			const bool spec = !(getenv("HOSTSIM_RESCUE_SPEC") && !atoi(getenv("HOSTSIM_RESCUE_SPEC")));
			std::vector<RTask> tasks; std::vector<LocalRes> res; std::vector<int> t_base(n_pairs + 1, 0);
			if (spec) {
				for (int p = 0; p < n_pairs; ++p) t_base[p + 1] = t_base[p] + rescue_enum(V, p, 0, 0, win_cap);
				tasks.resize(t_base[n_pairs] + 1); res.resize(t_base[n_pairs] + 1);
				for (int p = 0; p < n_pairs; ++p) if (t_base[p + 1] > t_base[p]) rescue_enum(V, p, tasks.data() + t_base[p], (u32)p, win_cap);
				std::vector<uint8_t> tq(max_len + 16);
				for (int k = 0; k < t_base[n_pairs]; ++k) { // one task = what the SW kernel gives a warp
					const RTask &t = tasks[k];
					const int p = (int)t.slot, i = (int)(t.key >> 16), r = (int)(t.key & 3), is_rev = (r >> 1 != (r & 1));
					const uint8_t *ms = V.tc.seq + V.tc.read_off[2 * p + !i];
					for (int x = 0; x < t.l_ms; ++x) tq[is_rev ? t.l_ms - 1 - x : x] = is_rev ? (ms[x] < 4 ? 3 - ms[x] : 4) : ms[x];
					for (int x = 0; x < t.tlen; ++x) ref[x] = (uint8_t)ref_base(V.ix, t.rb + x);
					res[k] = sw_local(opt, t.l_ms, tq.data(), t.tlen, ref.data(), rescue_xtra(opt, t.l_ms), M.L);
				}
			}
			if (spec && getenv("HOSTSIM_RESCUE_DROP") && atoi(getenv("HOSTSIM_RESCUE_DROP")) > 0) { // tests: withhold every N-th result so that the replay's own computation runs
				const int nth = atoi(getenv("HOSTSIM_RESCUE_DROP"));
				std::vector<RTask> t2; std::vector<LocalRes> r2; std::vector<int> b2(n_pairs + 1, 0);
				for (int p = 0; p < n_pairs; ++p) {
					for (int k = t_base[p]; k < t_base[p + 1]; ++k) if (k % nth != 0) { t2.push_back(tasks[k]); r2.push_back(res[k]); }
					b2[p + 1] = (int)t2.size();
					if (b2[p + 1] == b2[p] && t_base[p + 1] > t_base[p]) { t2.push_back(tasks[t_base[p]]); t2.back().key = 0xffffffffu; r2.push_back(res[t_base[p]]); b2[p + 1] = (int)t2.size(); } // keep the pair in the replay
				}
				t2.resize(t2.size() + 1); r2.resize(r2.size() + 1);
				tasks.swap(t2); res.swap(r2); t_base.swap(b2);
			}
			static unsigned int n_miss = 0;
			for (int p = 0; p < n_pairs; ++p) {
				if (spec) { // (a pair without a task cannot change: every replay step would be skipped or align nothing)
					if (t_base[p + 1] == t_base[p]) continue;
					RCache rc; rc.t = tasks.data() + t_base[p]; rc.res = res.data() + t_base[p]; rc.n = t_base[p + 1] - t_base[p]; rc.cur = 0; rc.miss = &n_miss;
					body_rescue(V, p, bbuf.data(), M, &rc);
				} else if (rescue_wanted(V, p)) body_rescue(V, p, bbuf.data(), M);
			}
			if (getenv("HOSTSIM_VERBOSE")) fprintf(stderr, "[hostsim] rescue: %d alignments computed ahead, %u computed in the replay\n", t_base[n_pairs], n_miss);
		}
		// planning
		std::vector<u64> tsoff(n + 1, 0);
		for (int i = 0; i < n; ++i) tsoff[i + 1] = tsoff[i] + 2ull * na[i] + 1;
		std::vector<PTask> tslots(tsoff[n] + 1); std::vector<ReadMeta> meta(n + 1);
		V.tslot_off = tsoff.data(); V.tslots = tslots.data(); V.meta = meta.data();
		for (int u = 0; u < n_units; ++u) body_plan(V, u);
		std::vector<u64> tkb(n + 1, 0);
		for (int i = 0; i < n; ++i) tkb[i + 1] = tkb[i] + meta[i].n_tasks;
		const u64 nt = tkb[n];
		std::vector<PTask> tasks(nt + 1); std::vector<AlnOut> outs(nt + 1); std::vector<u32> cigs((nt + 1) * CIG_CAP); std::vector<char> mds((nt + 1) * MD_CAP);
		V.tk_base = tkb.data(); V.tasks = tasks.data(); V.outs = outs.data(); V.cigs = cigs.data(); V.mds = mds.data();
		for (int r = 0; r < n; ++r) for (u32 i = 0; i < meta[r].n_tasks; ++i) tasks[tkb[r] + i] = tslots[tsoff[r] + i];
		A.g.z = zb.data(); A.g.zcap = 256 * 768;
		for (u64 t = 0; t < nt; ++t) body_cigar(V, t, A);
		// samblaster + dup-set
		std::vector<u64> k1(n_units + 1), k2(n_units + 1), smask(n + 1, 0); std::vector<uint8_t> valid(n_units + 1, 0), dup(n_units + 1, 0), disc(n_units + 1, 0);
		V.k1 = k1.data(); V.k2 = k2.data(); V.valid = valid.data(); V.dup = dup.data(); V.disc = disc.data(); V.split_mask = smask.data();
		if (sb.enabled) {
			for (int u = 0; u < n_units; ++u) body_sb(V, u);
			for (int u = 0; u < n_units; ++u) if (valid[u]) dup[u] = seen.insert(std::make_pair(k1[u], k2[u])).second ? 0 : 1;
		}
		// text
		std::vector<u64> len[3], toff[3];
		for (int k = 0; k < 3; ++k) { len[k].assign(n + 1, 0); toff[k].assign(n + 1, 0); V.len[k] = len[k].data(); V.off[k] = toff[k].data(); }
		for (int r = 0; r < n; ++r) body_text<false>(V, r);
		for (int k = 0; k < 3; ++k) { for (int r = 0; r < n; ++r) toff[k][r + 1] = toff[k][r] + len[k][r]; text[k].assign(toff[k][n], '\0'); V.text[k] = &text[k][0]; }
		for (int r = 0; r < n; ++r) body_text<true>(V, r);
		read_off = toff[0];
		if (want_bam) { // BAM records: sizes + keys per line, stable sort by (tid, pos, strand), offsets, bytes — the steps of ssq_pipe.cu
			std::vector<u64> line_base(n + 1, 0);
			for (int r = 0; r < n; ++r) line_base[r + 1] = line_base[r] + read_n_lines(V, r);
			const u64 nl = line_base[n];
			std::vector<u32> line_read(nl + 1), perm(nl + 1); std::vector<u64> key(nl + 1), bsz[3], boff[3];
			for (int k = 0; k < 3; ++k) { bsz[k].assign(nl + 1, 0); boff[k].assign(nl + 1, 0); V.bam_size[k] = bsz[k].data(); }
			V.line_base = line_base.data(); V.line_read = line_read.data(); V.bam_key = key.data(); V.bam_blank_side = bam_blank_side;
			for (int r = 0; r < n; ++r) body_bam_size(V, r);
			for (u64 i = 0; i < nl; ++i) perm[i] = (u32)i;
			std::stable_sort(perm.begin(), perm.begin() + nl, [&](u32 a, u32 b) { return key[a] < key[b]; });
			V.bam_perm = perm.data();
			for (int k = 0; k < 3; ++k) { u64 at = 0; for (u64 i = 0; i < nl; ++i) { boff[k][i] = at; at += bsz[k][perm[i]]; } bam[k].assign(at, '\0'); V.bam[k] = &bam[k][0]; V.bam_off[k] = boff[k].data(); }
			for (u64 i = 0; i < nl; ++i) body_bam_write(V, i);
		}
		return err;
	}
};
static HostPipe g_pipe;

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

// `bwa mem` for a batch of pairs through the product's host orchestration + the host backend -> malloc'd SAM text
static void pipe_index_names(const ssqo_idx_t *idx, std::vector<std::string> &ctg, std::vector<i32> &len)
{
	for (int i = 0; i < idx->bns.n_seqs; ++i) { ctg.push_back(idx->bns.anns[i].name); len.push_back(idx->bns.anns[i].len); }
}
// `bwa mem` records for a batch (plain: no samblaster stage) -> malloc'd SAM text
char *hostsim_mem_pe(const ssqo_idx_t *idx, int n_reads, const char **names, const char **seqs, const char **quals, int64_t n_processed, const char *rg_id, int paired)
{
	std::vector<i64> aoff_; std::vector<i32> alen_; DevIndex ix = make_ix(idx, aoff_, alen_);
	ssq_opts_t opt; ssq_opts_default(&opt);
	SbOpts sb; memset(&sb, 0, sizeof sb);
	std::vector<std::string> ctg; std::vector<i32> clen; pipe_index_names(idx, ctg, clen);
	HostPipe hp;
	if (hp.run(ix, opt, sb, n_reads, names, seqs, quals, 0, n_processed, paired, 0, rg_id, ctg, clen)) return 0;
	char *out = (char*)malloc(hp.text[0].size() + 1);
	memcpy(out, hp.text[0].c_str(), hp.text[0].size() + 1);
	return out;
}
// the fused `bwa mem | samblaster` pipeline for one batch; sbv = {excludeDups, addMateTags, maxSplitCount, minNonOverlap, removeDups};
// reset != 0 starts a new run (empties the dup-set); pes_in: optional 4 x {low, high, failed, avg, std} as doubles (-I)
int hostsim_pipe(const ssqo_idx_t *idx, int n_reads, const char **names, const char **seqs, const char **quals, const char **comments, int64_t n_processed, const char *rg_id, int paired,
                 const int *sbv, int reset, const double *pes_in, char **out_main, char **out_split, char **out_disc)
{
	std::vector<i64> aoff_; std::vector<i32> alen_; DevIndex ix = make_ix(idx, aoff_, alen_);
	ssq_opts_t opt; ssq_opts_default(&opt);
	SbOpts sb; memset(&sb, 0, sizeof sb);
	sb.enabled = 1; sb.excludeDups = sbv[0]; sb.addMateTags = sbv[1]; sb.maxSplitCount = sbv[2]; sb.minNonOverlap = sbv[3]; sb.removeDups = sbv[4]; sb.minIndelSize = 50; sb.maxUnmappedBases = 50; sb.want_split = sb.want_disc = 1;
	std::vector<std::string> ctg; std::vector<i32> clen; pipe_index_names(idx, ctg, clen);
	if (reset) g_pipe.seen.clear();
	PeStat pes[4]; memset(pes, 0, sizeof pes);
	if (pes_in) for (int d = 0; d < 4; ++d) { pes[d].low = (int)pes_in[5 * d]; pes[d].high = (int)pes_in[5 * d + 1]; pes[d].failed = (int)pes_in[5 * d + 2]; pes[d].avg = pes_in[5 * d + 3]; pes[d].std = pes_in[5 * d + 4]; }
	const int rc = g_pipe.run(ix, opt, sb, n_reads, names, seqs, quals, comments, n_processed, paired, pes_in ? pes : 0, rg_id, ctg, clen);
	if (rc) return rc;
	char **outs[3] = {out_main, out_split, out_disc};
	for (int k = 0; k < 3; ++k) { *outs[k] = (char*)malloc(g_pipe.text[k].size() + 1); memcpy(*outs[k], g_pipe.text[k].c_str(), g_pipe.text[k].size() + 1); }
	return 0;
}
// BAM records of the last hostsim_pipe_bam() batch; call with want = 1 before hostsim_pipe to switch the encoder on
void hostsim_pipe_want_bam(int want, int blank_side) { g_pipe.want_bam = want; g_pipe.bam_blank_side = blank_side; }
uint64_t hostsim_pipe_bam(int stream, char *out, uint64_t cap)
{
	const std::string &b = g_pipe.bam[stream];
	if (out && cap >= b.size()) memcpy(out, b.data(), b.size());
	return b.size();
}

// randomized check of ChainBuilder's tree-ordered chains against a plain ordered-array restatement (look-up = first chain with
// an equal pos else the predecessor; insert right after the slot): seed lists dense with equal and near-equal positions
int hostsim_chain_selftest(int n_trials, unsigned seed)
{
	DevIndex ix; memset(&ix, 0, sizeof ix);
	i64 aoff[1] = {0}; i32 alen[1] = {1 << 30};
	ix.l_pac = 1 << 30; ix.n_seqs = 1; ix.ann_off = aoff; ix.ann_len = alen;
	ssq_opts_t opt; ssq_opts_default(&opt);
	int bad = 0;
	srand(seed);
	for (int t = 0; t < n_trials; ++t) {
		const int n = 1 + rand() % 300;
		std::vector<Seed> seeds(n);
		for (int i = 0; i < n; ++i) {
			seeds[i].rbeg = (rand() % 40) * 50 + (rand() % 3 == 0 ? rand() % 200 : 0) + ((rand() & 1) ? 0 : (1ll << 30)); // both strands, many ties
			seeds[i].qbeg = rand() % 130; seeds[i].len = 19 + rand() % 30;
		}
		// reference: ordered array
		std::vector<int> chain_ref(n, -1), ord_ref; std::vector<ChainRec> chr(n); int nch = 0;
		for (int i = 0; i < n; ++i) {
			const Seed s = seeds[i];
			int lo = 0, hi = nch;
			while (lo < hi) { int mid = (lo + hi) >> 1; if (chr[ord_ref[mid]].pos < s.rbeg) lo = mid + 1; else hi = mid; }
			int slot = (lo < nch && chr[ord_ref[lo]].pos == s.rbeg) ? lo : lo - 1, merged = 0;
			if (nch && slot >= 0) {
				ChainRec &c = chr[ord_ref[slot]];
				i64 qend = c.last_q + c.last_len, rend = c.last_r + c.last_len;
				if (s.qbeg >= c.first_q && s.qbeg + s.len <= qend && s.rbeg >= c.first_r && s.rbeg + s.len <= rend) merged = 1;
				else if (!((c.last_r < ix.l_pac || c.first_r < ix.l_pac) && s.rbeg >= ix.l_pac)) {
					i64 x = s.qbeg - c.last_q, y = s.rbeg - c.last_r;
					if (y >= 0 && x - y <= opt.w && y - x <= opt.w && x - c.last_len < opt.max_chain_gap && y - c.last_len < opt.max_chain_gap) { c.last_q = s.qbeg; c.last_r = s.rbeg; c.last_len = s.len; ++c.n; chain_ref[i] = ord_ref[slot]; merged = 1; }
				}
			}
			if (!merged) {
				ChainRec &c = chr[nch]; c.pos = s.rbeg; c.first_r = c.last_r = s.rbeg; c.first_q = c.last_q = s.qbeg; c.last_len = s.len; c.rid = 0; c.n = 1;
				ord_ref.insert(ord_ref.begin() + (slot + 1), nch); chain_ref[i] = nch; ++nch;
			}
		}
		// ChainBuilder
		std::vector<i32> chain_of(n), ord(n); std::vector<ChainRec> ch(n), outc(n); std::vector<WIdx> wi(n); std::vector<Seed> sorted(n); std::vector<KeptChain> kp(n);
		ChainBuilder b; b.init(150, n, seeds.data(), 0, chain_of.data(), ch.data(), ord.data(), wi.data(), sorted.data(), outc.data(), kp.data());
		for (int i = 0; i < n; ++i) b.add_seed(ix, opt, i);
		b.inorder();
		bool ok = b.n_ch == nch;
		for (int i = 0; ok && i < n; ++i) ok = chain_of[i] == chain_ref[i];
		for (int k = 0; ok && k < nch; ++k) ok = ord[k] == ord_ref[k];
		bad += !ok;
	}
	return bad;
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
	o->mapQ_coef_len = 50; o->mapQ_coef_fac = 3; o->n_threads = 1;
}
}
