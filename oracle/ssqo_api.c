/*
 * ssqo_api.c — ORACLE (test infrastructure): flat, ctypes-friendly batch entry points that mirror
 * the product's C-ABI (include/ssq.h) one-to-one so that tests/ can compare array with array.
 * Every function here only forwards to the restatements in ssqo_bwt.c / ssqo_ksw.c / ssqo_mem.c.
 */
#include <stdlib.h>
#include <string.h>
#include "ssqo.h"
#include "ssqo_mem.h"
#include "ssqo_par.h"

/* mirrors ssq_smem_t / ssq_seed_t / ssq_sw_task_t / ssq_sw_result_t / ssq_alnreg_t of include/ssq.h */
typedef struct { uint64_t k, l, s; uint32_t qbeg, qend; } api_smem_t;
typedef struct { int64_t rbeg; int32_t qbeg, len; } api_seed_t;
typedef struct { uint64_t q_off, t_off; int32_t qlen, tlen, h0, w, end_bonus, zdrop; } api_sw_task_t;
typedef struct { int32_t score, qle, tle, gtle, gscore, max_off; } api_sw_result_t;
typedef struct {
	int64_t rb, re;
	int32_t qb, qe, rid, score, truesc, w, seedcov, seedlen0;
	float frac_rep;
	int32_t read_id;
} api_alnreg_t;

void ssqo_api_counters(ssqo_counters_t *out, int reset)
{
	if (out) *out = ssqo_cnt;
	if (reset) memset(&ssqo_cnt, 0, sizeof ssqo_cnt);
}

uint64_t ssqo_api_idx_info(const ssqo_idx_t *idx, int what)
{
	switch (what) {
	case 0: return (uint64_t)idx->bns.l_pac;
	case 1: return idx->bwt.seq_len;
	case 2: return idx->bwt.primary;
	case 3: return (uint64_t)idx->bns.n_seqs;
	case 4: return idx->bwt.bwt_size;
	case 5: return idx->bwt.n_sa;
	}
	return 0;
}

/* the three seeding passes for every read; out_off[i]..out_off[i+1] are read i's intervals (sorted by info) */
int64_t ssqo_api_smem_batch(const ssqo_idx_t *idx, int n_reads, const uint8_t *seq, const uint64_t *read_off,
                            api_smem_t *out, uint64_t out_cap, uint64_t *out_off)
{
	ssqo_opt_t opt;
	ssqo_intv_v v = {0, 0, 0};
	uint64_t n = 0;
	int i;
	size_t j;
	ssqo_opt_init(&opt);
	for (i = 0; i < n_reads; ++i) {
		int len = (int)(read_off[i + 1] - read_off[i]);
		out_off[i] = n;
		if (len < opt.min_seed_len) continue;
		ssqo_collect_intv(&opt, &idx->bwt, len, seq + read_off[i], &v);
		for (j = 0; j < v.n; ++j, ++n) {
			if (n >= out_cap) { free(v.a); return -1; }
			out[n].k = v.a[j].x[0]; out[n].l = v.a[j].x[1]; out[n].s = v.a[j].x[2];
			out[n].qbeg = (uint32_t)(v.a[j].info >> 32); out[n].qend = (uint32_t)v.a[j].info;
		}
	}
	out_off[n_reads] = n;
	free(v.a);
	return (int64_t)n;
}

int ssqo_api_sa_batch(const ssqo_idx_t *idx, uint64_t n, const uint64_t *rows, uint64_t *pos)
{
	uint64_t i;
	for (i = 0; i < n; ++i) pos[i] = ssqo_sa(&idx->bwt, rows[i]);
	return 0;
}

int ssqo_api_sw_extend_batch(uint64_t n, const api_sw_task_t *t, const uint8_t *qbuf, const uint8_t *tbuf, api_sw_result_t *r)
{
	ssqo_opt_t opt;
	uint64_t i;
	ssqo_opt_init(&opt);
	for (i = 0; i < n; ++i)
		r[i].score = ssqo_ksw_extend2(t[i].qlen, qbuf + t[i].q_off, t[i].tlen, tbuf + t[i].t_off, 5, opt.mat,
		                              opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, t[i].w, t[i].end_bonus, t[i].zdrop, t[i].h0,
		                              &r[i].qle, &r[i].tle, &r[i].gtle, &r[i].gscore, &r[i].max_off);
	return 0;
}

/* chains after mem_chain + mem_chain_flt, flattened: per read, chains in filter order; seeds in chain order */
int64_t ssqo_api_chain_batch(const ssqo_idx_t *idx, int n_reads, const uint8_t *seq, const uint64_t *read_off,
                             api_seed_t *seeds, uint64_t seed_cap, uint64_t *chain_seed_off /* [chain_cap+1] */, uint64_t chain_cap,
                             uint64_t *read_chain_off /* [n_reads+1] */)
{
	ssqo_opt_t opt;
	uint64_t ns = 0, nc = 0;
	int i, j, k;
	ssqo_opt_init(&opt);
	for (i = 0; i < n_reads; ++i) {
		int len = (int)(read_off[i + 1] - read_off[i]);
		ssqo_chain_v chn;
		read_chain_off[i] = nc;
		chn = ssqo_mem_chain(&opt, idx, len, seq + read_off[i]);
		chn.n = ssqo_chain_flt(&opt, (int)chn.n, chn.a);
		for (j = 0; j < (int)chn.n; ++j) {
			if (nc >= chain_cap) return -1;
			chain_seed_off[nc++] = ns;
			for (k = 0; k < chn.a[j].n; ++k, ++ns) {
				if (ns >= seed_cap) return -1;
				seeds[ns].rbeg = chn.a[j].seeds[k].rbeg; seeds[ns].qbeg = chn.a[j].seeds[k].qbeg; seeds[ns].len = chn.a[j].seeds[k].len;
			}
			free(chn.a[j].seeds);
		}
		free(chn.a);
	}
	chain_seed_off[nc] = ns;
	read_chain_off[n_reads] = nc;
	return (int64_t)nc;
}

typedef struct { const ssqo_opt_t *opt; const ssqo_idx_t *idx; const uint8_t *seq; const uint64_t *read_off; int stage; ssqo_alnreg_v *regs; } aworker_t;
static void worker_regs(void *d, long i, int tid)
{
	aworker_t *w = (aworker_t*)d;
	int len = (int)(w->read_off[i + 1] - w->read_off[i]), c;
	uint8_t *s = (uint8_t*)malloc(len + 1);
	ssqo_chain_v chn;
	(void)tid;
	memcpy(s, w->seq + w->read_off[i], len);
	chn = ssqo_mem_chain(w->opt, w->idx, len, s);
	chn.n = ssqo_chain_flt(w->opt, (int)chn.n, chn.a);
	for (c = 0; c < (int)chn.n; ++c) { ssqo_chain2aln(w->opt, w->idx, len, s, &chn.a[c], &w->regs[i]); free(chn.a[c].seeds); }
	free(chn.a);
	if (w->stage >= 1) w->regs[i].n = ssqo_sort_dedup_patch(w->opt, w->idx, s, (int)w->regs[i].n, w->regs[i].a);
	free(s);
}

/* alignment regions per read: stage 0 = straight out of seed extension (chain2aln order), 1 = after sort/dedup/patch */
int64_t ssqo_api_align_batch(const ssqo_idx_t *idx, int n_reads, const uint8_t *seq, const uint64_t *read_off, int stage,
                             api_alnreg_t *out, uint64_t out_cap, uint64_t *out_off, int n_threads)
{
	ssqo_opt_t opt;
	ssqo_alnreg_v *regs = (ssqo_alnreg_v*)calloc(n_reads, sizeof(ssqo_alnreg_v));
	uint64_t n = 0;
	int i;
	size_t j;
	ssqo_opt_init(&opt);
	if (n_threads < 1) n_threads = 1;
	{
		aworker_t w;
		w.opt = &opt; w.idx = idx; w.seq = seq; w.read_off = read_off; w.stage = stage; w.regs = regs;
		ssqo_parallel_for(n_threads, n_reads, worker_regs, &w);
	}
	for (i = 0; i < n_reads; ++i) {
		out_off[i] = n;
		for (j = 0; j < regs[i].n; ++j, ++n) {
			const ssqo_alnreg_t *p = &regs[i].a[j];
			if (out && n < out_cap) {
				api_alnreg_t *q = &out[n];
				q->rb = p->rb; q->re = p->re; q->qb = p->qb; q->qe = p->qe; q->rid = p->rid; q->score = p->score;
				q->truesc = p->truesc; q->w = p->w; q->seedcov = p->seedcov; q->seedlen0 = p->seedlen0;
				q->frac_rep = p->frac_rep; q->read_id = i;
			}
		}
		free(regs[i].a);
	}
	out_off[n_reads] = n;
	free(regs);
	return (int64_t)n;
}

/* `bwa mem` on an in-memory batch of paired reads -> concatenated SAM text (caller frees with ssqo_api_free) */
char *ssqo_api_mem_pe(const ssqo_idx_t *idx, int n_reads, const char **names, const char **seqs, const char **quals,
                      int64_t n_processed, int n_threads, const char *rg_id)
{
	ssqo_opt_t opt;
	ssqo_read_t *r = (ssqo_read_t*)calloc(n_reads, sizeof(ssqo_read_t));
	size_t l = 0, m = 1 << 16;
	char *out = (char*)malloc(m);
	int i;
	ssqo_opt_init(&opt);
	opt.flag |= SSQO_F_PE; opt.n_threads = n_threads > 0 ? n_threads : 1;
	for (i = 0; i < n_reads; ++i) {
		r[i].name = strdup(names[i]); r[i].seq = strdup(seqs[i]); r[i].qual = quals && quals[i] ? strdup(quals[i]) : 0;
		r[i].l_seq = (int)strlen(seqs[i]); r[i].id = i;
	}
	ssqo_process_seqs(&opt, idx, n_processed, n_reads, r, 0, rg_id);
	out[0] = 0;
	for (i = 0; i < n_reads; ++i) {
		size_t k = r[i].sam ? strlen(r[i].sam) : 0;
		if (l + k + 1 > m) { while (l + k + 1 > m) m <<= 1; out = (char*)realloc(out, m); }
		if (k) memcpy(out + l, r[i].sam, k);
		l += k; out[l] = 0;
		free(r[i].name); free(r[i].seq); free(r[i].qual); free(r[i].sam);
	}
	free(r);
	return out;
}

void ssqo_api_free(void *p) { free(p); }
