// ssq_mem_host.h — host orchestration of `bwa mem` for one batch, templated on the backend that executes the
// base-level stages.  The product instantiates it with the CUDA backend (ssq_mem.cu: kernels k_dedup / k_matesw /
// k_cigar after ssq_batch_run); tests/hostsim instantiates it with a backend that calls the same SSQ_HD routines on the
// host, so that the SAM text can be diffed against the oracle on a box without a GPU.  Nothing in here touches bases or
// DP cells: insert-size statistics, primary marking, pairing, MAPQ (IEEE doubles, libm erfc/log like the reference) and
// SAM formatting.  Upstream routines replaced: mem_pestat, mem_mark_primary_se, mem_approx_mapq_se, mem_pair,
// mem_sam_pe, mem_gen_alt, mem_reg2sam, mem_aln2sam (inside `$BWA mem`, /root/reference/bin/speedseq:438,468).
#pragma once
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <thread>
#include <chrono>
#include <vector>
#include "ssq_dev2.cuh"

#define CIG_CAP 64
#define MD_CAP 512
struct CigTask { AlnReg reg; i32 read, pad; };
struct HostIndexInfo { i64 l_pac; int n_seqs; char **names; const i64 *ann_off; };

static inline u64 hash64(u64 key)
{
	key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
	key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
	return key;
}


// ---- insert-size statistics of one batch (mem_pestat) ----
static int cal_sub(const ssq_opts_t &o, const AlnReg *r, int n)
{
	int j;
	for (j = 1; j < n; ++j) {
		const int b_max = r[j].qb > r[0].qb ? r[j].qb : r[0].qb, e_min = r[j].qe < r[0].qe ? r[j].qe : r[0].qe;
		if (e_min > b_max) {
			const int min_l = r[j].qe - r[j].qb < r[0].qe - r[0].qb ? r[j].qe - r[j].qb : r[0].qe - r[0].qb;
			if (e_min - b_max >= min_l * o.mask_level) break;
		}
	}
	return j < n ? r[j].score : o.min_seed_len * o.a;
}

struct U64Lt { bool operator()(u64 a, u64 b) const { return a < b; } };

static void pestat(const ssq_opts_t &o, i64 l_pac, int n_reads, const AlnReg *areg, const u64 *areg_off, const u32 *n_areg, PeStat pes[4], FILE *log)
{
	std::vector<u64> isize[4];
	memset(pes, 0, 4 * sizeof(PeStat));
	for (int i = 0; i < n_reads >> 1; ++i) {
		const AlnReg *r0 = areg + areg_off[2 * i], *r1 = areg + areg_off[2 * i + 1];
		const int n0 = (int)n_areg[2 * i], n1 = (int)n_areg[2 * i + 1];
		i64 is;
		if (n0 == 0 || n1 == 0) continue;
		if (cal_sub(o, r0, n0) > 0.8 * r0[0].score) continue;
		if (cal_sub(o, r1, n1) > 0.8 * r1[0].score) continue;
		if (r0[0].rid != r1[0].rid) continue;
		const int dir = infer_dir(l_pac, r0[0].rb, r1[0].rb, &is);
		if (is && is <= o.max_ins) isize[dir].push_back((u64)is);
	}
	if (log) fprintf(log, "[M::mem_pestat] # candidate unique pairs for (FF, FR, RF, RR): (%ld, %ld, %ld, %ld)\n", (long)isize[0].size(), (long)isize[1].size(), (long)isize[2].size(), (long)isize[3].size());
	for (int d = 0; d < 4; ++d) {
		PeStat *r = &pes[d];
		std::vector<u64> &q = isize[d];
		if (q.size() < 10) { if (log) fprintf(log, "[M::mem_pestat] skip orientation %c%c as there are not enough pairs\n", "FR"[d >> 1 & 1], "FR"[d & 1]); r->failed = 1; continue; }
		if (log) fprintf(log, "[M::mem_pestat] analyzing insert size distribution for orientation %c%c...\n", "FR"[d >> 1 & 1], "FR"[d & 1]);
		ks_introsort((long)q.size(), q.data(), U64Lt());
		const int p25 = (int)q[(int)(.25 * q.size() + .499)], p50 = (int)q[(int)(.50 * q.size() + .499)], p75 = (int)q[(int)(.75 * q.size() + .499)];
		int x = 0;
		r->low = (int)(p25 - 2.0 * (p75 - p25) + .499);
		if (r->low < 1) r->low = 1;
		r->high = (int)(p75 + 2.0 * (p75 - p25) + .499);
		if (log) fprintf(log, "[M::mem_pestat] (25, 50, 75) percentile: (%d, %d, %d)\n[M::mem_pestat] low and high boundaries for computing mean and std.dev: (%d, %d)\n", p25, p50, p75, r->low, r->high);
		r->avg = 0;
		for (size_t i = 0; i < q.size(); ++i) if (q[i] >= (u64)r->low && q[i] <= (u64)r->high) { r->avg += q[i]; ++x; }
		r->avg /= x;
		r->std = 0;
		for (size_t i = 0; i < q.size(); ++i) if (q[i] >= (u64)r->low && q[i] <= (u64)r->high) r->std += (q[i] - r->avg) * (q[i] - r->avg);
		r->std = sqrt(r->std / x);
		if (log) fprintf(log, "[M::mem_pestat] mean and std.dev: (%.2f, %.2f)\n", r->avg, r->std);
		r->low = (int)(p25 - 3.0 * (p75 - p25) + .499);
		r->high = (int)(p75 + 3.0 * (p75 - p25) + .499);
		if (r->low > r->avg - 4.0 * r->std) r->low = (int)(r->avg - 4.0 * r->std + .499);
		if (r->high < r->avg + 4.0 * r->std) r->high = (int)(r->avg + 4.0 * r->std + .499);
		if (r->low < 1) r->low = 1;
		if (log) fprintf(log, "[M::mem_pestat] low and high boundaries for proper pairs: (%d, %d)\n", r->low, r->high);
	}
	size_t max = 0;
	for (int d = 0; d < 4; ++d) max = max > isize[d].size() ? max : isize[d].size();
	for (int d = 0; d < 4; ++d)
		if (pes[d].failed == 0 && isize[d].size() < max * 0.05) { pes[d].failed = 1; if (log) fprintf(log, "[M::mem_pestat] skip orientation %c%c\n", "FR"[d >> 1 & 1], "FR"[d & 1]); }
}

// ---- primary marking / MAPQ / pairing ----
static int mark_primary(const ssq_opts_t &o, int n, AlnReg *a, i64 id)
{
	if (n == 0) return 0;
	for (int i = 0; i < n; ++i) { a[i].sub = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = hash64((u64)(id + i)); }
	ks_introsort((long)n, a, ArsHashLt());
	int tmp = o.a + o.b;
	tmp = o.o_del + o.e_del > tmp ? o.o_del + o.e_del : tmp;
	tmp = o.o_ins + o.e_ins > tmp ? o.o_ins + o.e_ins : tmp;
	std::vector<int> z;
	z.push_back(0);
	for (int i = 1; i < n; ++i) {
		size_t k;
		for (k = 0; k < z.size(); ++k) {
			const int j = z[k];
			const int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb, e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
			if (e_min > b_max) {
				const int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
				if (e_min - b_max >= min_l * o.mask_level) {
					if (a[j].sub == 0) a[j].sub = a[i].score;
					if (a[j].score - a[i].score <= tmp) ++a[j].sub_n;
					break;
				}
			}
		}
		if (k == z.size()) z.push_back(i); else a[i].secondary = z[k];
	}
	for (int i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
	return n;
}

static int approx_mapq_se(const ssq_opts_t &o, const AlnReg &a)
{
	int mapq, l, sub = a.sub ? a.sub : o.min_seed_len * o.a;
	sub = a.csub > sub ? a.csub : sub;
	if (sub >= a.score) return 0;
	l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
	const double identity = 1. - (double)(l * o.a - a.score) / (o.a + o.b) / l;
	if (a.score == 0) mapq = 0;
	else {
		double t = l < o.mapQ_coef_len ? 1. : o.mapQ_coef_fac / log(l);
		t *= identity * identity;
		mapq = (int)(6.02 * (a.score - sub) / o.a * t * t + .499);
	}
	if (a.sub_n > 0) mapq -= (int)(4.343 * log(a.sub_n + 1) + .499);
	if (mapq > 60) mapq = 60;
	if (mapq < 0) mapq = 0;
	return (int)(mapq * (1. - a.frac_rep) + .499);
}
#define raw_mapq(diff, a) ((int)(6.02 * (diff) / (a) + .499))

struct P64 { u64 x, y; };
struct P64Lt { bool operator()(const P64 &a, const P64 &b) const { return a.x < b.x || (a.x == b.x && a.y < b.y); } };

static int mem_pair(const ssq_opts_t &o, const HostIndexInfo *ix, const PeStat pes[4], AlnReg *const a[2], const int n[2], i64 id, int *sub, int *n_sub, int z[2])
{
	const i64 l_pac = ix->l_pac;
	std::vector<P64> v, u;
	int y[4], ret;
	for (int r = 0; r < 2; ++r)
		for (int i = 0; i < n[r]; ++i) {
			const AlnReg &e = a[r][i];
			P64 key;
			key.x = e.rb < l_pac ? e.rb : (l_pac << 1) - 1 - e.rb;
			key.x = (u64)e.rid << 32 | (key.x - ix->ann_off[e.rid]);
			key.y = (u64)e.score << 32 | i << 2 | (e.rb >= l_pac) << 1 | r;
			v.push_back(key);
		}
	ks_introsort((long)v.size(), v.data(), P64Lt());
	y[0] = y[1] = y[2] = y[3] = -1;
	for (int i = 0; i < (int)v.size(); ++i) {
		for (int r = 0; r < 2; ++r) {
			const int dir = r << 1 | (int)(v[i].y >> 1 & 1);
			if (pes[dir].failed) continue;
			const int which = r << 1 | (int)((v[i].y & 1) ^ 1);
			if (y[which] < 0) continue;
			for (int k = y[which]; k >= 0; --k) {
				if ((int)(v[k].y & 3) != which) continue;
				const i64 dist = (i64)v[i].x - (i64)v[k].x;
				if (dist > pes[dir].high) break;
				if (dist < pes[dir].low) continue;
				const double ns = (dist - pes[dir].avg) / pes[dir].std;
				int q = (int)((v[i].y >> 32) + (v[k].y >> 32) + .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * o.a + .499);
				if (q < 0) q = 0;
				P64 p;
				p.y = (u64)k << 32 | (u32)i;
				p.x = (u64)q << 32 | (hash64(p.y ^ (u64)id << 8) & 0xffffffffU);
				u.push_back(p);
			}
		}
		y[v[i].y & 3] = i;
	}
	if (!u.empty()) {
		int tmp = o.a + o.b;
		tmp = tmp > o.o_del + o.e_del ? tmp : o.o_del + o.e_del;
		tmp = tmp > o.o_ins + o.e_ins ? tmp : o.o_ins + o.e_ins;
		ks_introsort((long)u.size(), u.data(), P64Lt());
		const int i = (int)(u.back().y >> 32), k = (int)(u.back().y << 32 >> 32);
		z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
		z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
		ret = (int)(u.back().x >> 32);
		*sub = u.size() > 1 ? (int)(u[u.size() - 2].x >> 32) : 0;
		*n_sub = 0;
		for (int j = (int)u.size() - 2; j >= 0; --j) if (*sub - (int)(u[j].x >> 32) <= tmp) ++*n_sub;
	} else { ret = 0; *sub = 0; *n_sub = 0; }
	return ret;
}

// ---- SAM assembly ----
struct Aln { // one output alignment (host): AlnOut + CIGAR + MD + XA text
	i64 pos; int rid, flag, is_rev, mapq, NM, score, sub; bool has;
	std::vector<u32> cigar; std::string md, xa;
	Aln() : pos(-1), rid(-1), flag(0), is_rev(0), mapq(0), NM(0), score(0), sub(0), has(false) {}
};

struct PendingAln { int read, slot; }; // which Aln a CIGAR task fills

struct ReadPlan { // what is written for one read
	std::vector<Aln> lines; // primary first, then supplementary
	int extra_flag;
	std::vector<std::string> xa_for;
	ReadPlan() : extra_flag(0) {} // XA text per region index (built from xa tasks)
};

static inline void put_num(std::string &s, long long v) // == "%lld"
{
	char b[24]; int n = 24;
	unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
	do { b[--n] = (char)('0' + u % 10); u /= 10; } while (u);
	if (v < 0) b[--n] = '-';
	s.append(b + n, 24 - n);
}
static int get_rlen(const std::vector<u32> &c) { int l = 0; for (size_t k = 0; k < c.size(); ++k) { int op = c[k] & 0xf; if (op == 0 || op == 2) l += c[k] >> 4; } return l; }

static void aln2sam(const HostIndexInfo *ix, std::string &str, const char *name, const char *seq_codes, int l_seq, const char *qual, const std::vector<Aln> &list, int which,
                    const Aln *m_, const char *rg_id, const char *comment)
{
	// mem_aln2sam() works on copies of the alignment and its mate and patches them (an unmapped end takes the other end's
	// coordinates and loses its CIGAR); here the patched scalars live in locals and the records themselves are never copied
	static const std::vector<u32> no_cigar;
	const Aln &p0 = list[which];
	struct { i64 pos; int rid, flag, is_rev, mapq, NM, score, sub; const std::vector<u32> &cigar_ref() const { return *cig; } const std::vector<u32> *cig; const std::string *mdp, *xap; } pv;
	pv.pos = p0.pos; pv.rid = p0.rid; pv.flag = p0.flag; pv.is_rev = p0.is_rev; pv.mapq = p0.mapq; pv.NM = p0.NM; pv.score = p0.score; pv.sub = p0.sub;
	pv.cig = &p0.cigar; pv.mdp = &p0.md; pv.xap = &p0.xa;
	struct { i64 pos; int rid, is_rev; const std::vector<u32> *cig; } mv;
	const bool has_m = m_ != 0;
	mv.pos = has_m ? m_->pos : -1; mv.rid = has_m ? m_->rid : -1; mv.is_rev = has_m ? m_->is_rev : 0; mv.cig = has_m ? &m_->cigar : &no_cigar;
	pv.flag |= has_m ? 0x1 : 0;
	pv.flag |= pv.rid < 0 ? 0x4 : 0;
	pv.flag |= has_m && mv.rid < 0 ? 0x8 : 0;
	if (pv.rid < 0 && has_m && mv.rid >= 0) { pv.rid = mv.rid; pv.pos = mv.pos; pv.is_rev = mv.is_rev; pv.cig = &no_cigar; }
	if (has_m && mv.rid < 0 && pv.rid >= 0) { mv.rid = pv.rid; mv.pos = pv.pos; mv.is_rev = pv.is_rev; mv.cig = &no_cigar; }
	pv.flag |= pv.is_rev ? 0x10 : 0;
	pv.flag |= has_m && mv.is_rev ? 0x20 : 0;
	struct PView { i64 pos; int rid, flag, is_rev, mapq, NM, score, sub; const std::vector<u32> &cigar; const std::string &md, &xa; };
	struct MView { i64 pos; int rid, is_rev; const std::vector<u32> &cigar; };
	const PView p = {pv.pos, pv.rid, pv.flag, pv.is_rev, pv.mapq, pv.NM, pv.score, pv.sub, *pv.cig, *pv.mdp, *pv.xap};
	const MView mview = {mv.pos, mv.rid, mv.is_rev, *mv.cig};
	const MView *m = has_m ? &mview : 0;
	str += name; str += '\t';
	put_num(str, (p.flag & 0xffff) | (p.flag & 0x10000 ? 0x100 : 0)); str += '\t';
	if (p.rid >= 0) {
		str += ix->names[p.rid]; str += '\t';
		put_num(str, p.pos + 1); str += '\t';
		put_num(str, p.mapq); str += '\t';
		if (!p.cigar.empty()) {
			for (size_t i = 0; i < p.cigar.size(); ++i) {
				int c = p.cigar[i] & 0xf;
				if (c == 3 || c == 4) c = which ? 4 : 3;
				put_num(str, p.cigar[i] >> 4); str += "MIDSH"[c];
			}
		} else str += '*';
	} else str += "*\t0\t0\t*";
	str += '\t';
	if (m && m->rid >= 0) {
		if (p.rid == m->rid) str += '='; else str += ix->names[m->rid];
		str += '\t';
		put_num(str, m->pos + 1); str += '\t';
		if (p.rid == m->rid) {
			const i64 p0 = p.pos + (p.is_rev ? get_rlen(p.cigar) - 1 : 0), p1 = m->pos + (m->is_rev ? get_rlen(m->cigar) - 1 : 0);
			if (m->cigar.empty() || p.cigar.empty()) str += '0';
			else put_num(str, -(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
		} else str += '0';
	} else str += "*\t0\t0";
	str += '\t';
	if (p.flag & 0x100) str += "*\t*";
	else {
		int qb = 0, qe = l_seq;
		if (!p.cigar.empty() && which) {
			const int c0 = p.cigar[0] & 0xf, c1 = p.cigar.back() & 0xf;
			if (!p.is_rev) { if (c0 == 4 || c0 == 3) qb += p.cigar[0] >> 4; if (c1 == 4 || c1 == 3) qe -= p.cigar.back() >> 4; }
			else { if (c0 == 4 || c0 == 3) qe -= p.cigar[0] >> 4; if (c1 == 4 || c1 == 3) qb += p.cigar.back() >> 4; }
		}
		char buf[2 * SSQ_MAX_READ_LEN + 8];
		int nb = 0;
		if (!p.is_rev) {
			for (int i = qb; i < qe; ++i) buf[nb++] = "ACGTN"[(int)seq_codes[i]];
			buf[nb++] = '\t';
			str.append(buf, nb);
			if (qual) str.append(qual + qb, qe - qb); else str += '*';
		} else {
			for (int i = qe - 1; i >= qb; --i) buf[nb++] = "TGCAN"[(int)seq_codes[i]];
			buf[nb++] = '\t';
			if (qual) for (int i = qe - 1; i >= qb; --i) buf[nb++] = qual[i]; else buf[nb++] = '*';
			str.append(buf, nb);
		}
	}
	if (!p.cigar.empty()) { str += "\tNM:i:"; put_num(str, p.NM); str += "\tMD:Z:"; str += p.md; }
	if (p.score >= 0) { str += "\tAS:i:"; put_num(str, p.score); }
	if (p.sub >= 0) { str += "\tXS:i:"; put_num(str, p.sub); }
	if (rg_id && rg_id[0]) { str += "\tRG:Z:"; str += rg_id; }
	if (!(p.flag & 0x100)) {
		size_t i;
		for (i = 0; i < list.size(); ++i) if ((int)i != which && !(list[i].flag & 0x100)) break;
		if (i < list.size()) {
			str += "\tSA:Z:";
			for (i = 0; i < list.size(); ++i) {
				const Aln &r = list[i];
				if ((int)i == which || (r.flag & 0x100)) continue;
				str += ix->names[r.rid]; str += ',';
				put_num(str, r.pos + 1); str += ',';
				str += "+-"[r.is_rev]; str += ',';
				for (size_t k = 0; k < r.cigar.size(); ++k) { put_num(str, r.cigar[k] >> 4); str += "MIDSH"[r.cigar[k] & 0xf]; }
				str += ','; put_num(str, r.mapq);
				str += ','; put_num(str, r.NM);
				str += ';';
			}
		}
	}
	if (!p.xa.empty()) { str += "\tXA:Z:"; str += p.xa; }
	if (comment) { str += '\t'; str += comment; }
	str += '\n';
}

// ---- the batch driver --------------------------------------------------------------------------------------------
struct CigReq { int read, reg_idx; AlnReg reg; int kind; /* 0 = output line, 1 = XA entry, 2 = mate header */ int line; int xa_owner; };

// Backend concept:
//   int  align(int n_reads, const uint8_t *codes, const u64 *off, int paired, int max_matesw)  seeding..extension + sort/dedup/patch; fills aoff/na/areg
//   int  rescue(const PeStat pes[4])                                                            mate rescue over all pairs; refreshes na/areg
//   int  cigar(const std::vector<CigTask>&, std::vector<AlnOut>&, std::vector<u32>&, std::vector<char>&)
//   std::vector<u64> aoff; std::vector<u32> na; std::vector<AlnReg> areg;                      region lists of every read (host copies)
template <class Backend>
static int mem_batch_sam(Backend &be, const ssq_opts_t &o, const HostIndexInfo *idx, int n_reads, const char *const *names, const uint8_t *codes_, const u64 *off_,
                         const char *const *quals, const char *const *comments, int64_t n_processed, int paired, const PeStat *pes0, const char *rg_id, FILE *logfp,
                         std::string &sam, std::string &err, std::vector<size_t> *line_off = 0)
{
	const std::vector<uint8_t> codes(codes_, codes_ + off_[n_reads] + 1);
	const std::vector<u64> off(off_, off_ + n_reads + 1);
	int rc;
	const bool timing = getenv("SSQ_MEM_TIMING") != 0;
	auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t_plan = 0, t_dist = 0, t_fmt = 0, t0 = 0;
	if ((rc = be.align(n_reads, codes_, off_, paired, o.max_matesw))) return rc;
	PeStat pes[4];
	memset(pes, 0, sizeof pes);
	if (paired && n_reads) {
		if (pes0) for (int d = 0; d < 4; ++d) pes[d] = pes0[d];
		else pestat(o, idx->l_pac, n_reads, be.areg.data(), be.aoff.data(), be.na.data(), pes, logfp);
		if ((rc = be.rescue(pes))) return rc;
	}
	std::vector<u64> &aoff = be.aoff; std::vector<u32> &na = be.na; std::vector<AlnReg> &areg = be.areg;
	// 5. host: primary marking, pairing, MAPQ; collect the alignments that need a CIGAR
	std::vector<CigReq> reqs;
	const int n_thr = o.n_threads > 1 ? (o.n_threads < 64 ? o.n_threads : 64) : 1;
	std::vector<std::vector<CigReq> > reqs_t(n_thr);
	std::vector<ReadPlan> plan(n_reads);
	std::vector<std::vector<Aln> > mate_hdr(n_reads); // h[] of the pair (mate info), index 0 only
	std::vector<int> z0(n_reads, -1);
	struct PairInfo { int mode; /* 0 = paired output, 1 = no_pairing */ int z[2], q_se[2], extra_flag; };
	std::vector<PairInfo> pinfo(paired ? n_reads / 2 : 0);
	auto add_req = [&](std::vector<CigReq> &reqs, int read, int reg_idx, int kind, int line, int xa_owner) { CigReq r; r.read = read; r.reg_idx = reg_idx; r.reg = areg[aoff[read] + reg_idx]; r.kind = kind; r.line = line; r.xa_owner = xa_owner; reqs.push_back(r); };
	auto plan_xa = [&](std::vector<CigReq> &reqs, int read) { // mem_gen_alt: secondary hits within 0.8x of their primary, at most max_XA_hits per primary
		AlnReg *a = areg.data() + aoff[read];
		const int n = (int)na[read];
		std::vector<int> cnt(n, 0);
		for (int i = 0; i < n; ++i) { const int k = a[i].secondary_all; if (k >= 0 && a[i].score >= a[k].score * (double)o.XA_drop_ratio) ++cnt[k]; }
		for (int i = 0; i < n; ++i) {
			const int k = a[i].secondary_all;
			if (!(k >= 0 && a[i].score >= a[k].score * (double)o.XA_drop_ratio)) continue;
			if (cnt[k] > o.max_XA_hits) continue;
			add_req(reqs, read, i, 1, -1, k);
		}
	};
	auto plan_reg2sam = [&](std::vector<CigReq> &reqs, int read, int extra_flag) { // mem_reg2sam without -a: every non-secondary hit above T, first = primary, rest supplementary
		AlnReg *a = areg.data() + aoff[read];
		const int n = (int)na[read];
		plan_xa(reqs, read);
		plan[read].extra_flag = extra_flag;
		int l = 0;
		for (int k = 0; k < n; ++k) {
			if (a[k].score < o.T || a[k].secondary >= 0) continue;
			plan[read].lines.push_back(Aln());
			Aln &q = plan[read].lines.back();
			q.has = true; q.flag = extra_flag | (l ? 0x800 : 0);
			q.mapq = approx_mapq_se(o, a[k]);
			if (l && q.mapq > plan[read].lines[0].mapq) q.mapq = plan[read].lines[0].mapq;
			add_req(reqs, read, k, 0, l, k);
			++l;
		}
	};
	auto plan_range = [&](int tid, int lo, int hi) { // units: reads (single-end) or pairs (paired); every unit only touches its own reads
	std::vector<CigReq> &reqs = reqs_t[tid];
	if (!paired) {
		for (int r = lo; r < hi; ++r) { mark_primary(o, (int)na[r], areg.data() + aoff[r], n_processed + r); plan_reg2sam(reqs, r, 0); }
	} else {
		for (int p = lo; p < hi; ++p) {
			AlnReg *a[2] = {areg.data() + aoff[2 * p], areg.data() + aoff[2 * p + 1]};
			int n[2] = {(int)na[2 * p], (int)na[2 * p + 1]}, z[2] = {0, 0}, osc = 0, subo = 0, n_sub = 0, extra_flag = 1;
			const i64 id = (n_processed >> 1) + p;
			PairInfo &pi = pinfo[p];
			mark_primary(o, n[0], a[0], id << 1 | 0);
			mark_primary(o, n[1], a[1], id << 1 | 1);
			bool pairing = false;
			if (n[0] && n[1] && (osc = mem_pair(o, idx, pes, a, n, id, &subo, &n_sub, z)) > 0) {
				bool multi = false;
				for (int i = 0; i < 2; ++i) for (int j = 1; j < n[i]; ++j) if (a[i][j].secondary < 0 && a[i][j].score >= o.T) { multi = true; break; }
				if (!multi) {
					int q_pe, score_un, q_se[2];
					pairing = true;
					score_un = a[0][0].score + a[1][0].score - o.pen_unpaired;
					subo = subo > score_un ? subo : score_un;
					q_pe = raw_mapq(osc - subo, o.a);
					if (n_sub > 0) q_pe -= (int)(4.343 * log(n_sub + 1) + .499);
					if (q_pe < 0) q_pe = 0;
					if (q_pe > 60) q_pe = 60;
					q_pe = (int)(q_pe * (1. - .5 * (a[0][0].frac_rep + a[1][0].frac_rep)) + .499);
					if (osc > score_un) {
						AlnReg *c[2] = {&a[0][z[0]], &a[1][z[1]]};
						for (int i = 0; i < 2; ++i) {
							if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
							q_se[i] = approx_mapq_se(o, *c[i]);
						}
						q_se[0] = q_se[0] > q_pe ? q_se[0] : q_pe < q_se[0] + 40 ? q_pe : q_se[0] + 40;
						q_se[1] = q_se[1] > q_pe ? q_se[1] : q_pe < q_se[1] + 40 ? q_pe : q_se[1] + 40;
						extra_flag |= 2;
						q_se[0] = q_se[0] < raw_mapq(c[0]->score - c[0]->csub, o.a) ? q_se[0] : raw_mapq(c[0]->score - c[0]->csub, o.a);
						q_se[1] = q_se[1] < raw_mapq(c[1]->score - c[1]->csub, o.a) ? q_se[1] : raw_mapq(c[1]->score - c[1]->csub, o.a);
					} else { z[0] = z[1] = 0; q_se[0] = approx_mapq_se(o, a[0][0]); q_se[1] = approx_mapq_se(o, a[1][0]); }
					for (int i = 0; i < 2; ++i) {
						const int k = a[i][z[i]].secondary_all;
						if (k >= 0 && k < n[i]) {
							for (int j = 0; j < n[i]; ++j) if (a[i][j].secondary_all == k || j == k) a[i][j].secondary_all = z[i];
							a[i][z[i]].secondary_all = -1;
						}
					}
					pi.mode = 0; pi.z[0] = z[0]; pi.z[1] = z[1]; pi.q_se[0] = q_se[0]; pi.q_se[1] = q_se[1]; pi.extra_flag = extra_flag;
					for (int i = 0; i < 2; ++i) {
						const int read = 2 * p + i;
						plan_xa(reqs, read);
						plan[read].lines.push_back(Aln());
						Aln &h = plan[read].lines.back();
						h.has = true; h.mapq = q_se[i]; h.flag = 0x40 << i | extra_flag;
						// reg2aln sets the secondary flag from the region: requests carry the (possibly updated) region
						add_req(reqs, read, z[i], 0, 0, z[i]);
					}
				}
			}
			if (!pairing) {
				pi.mode = 1; pi.extra_flag = 1;
				for (int i = 0; i < 2; ++i) { // mate header h[i]: the best hit if above T, else unmapped
					const int read = 2 * p + i;
					mate_hdr[read].push_back(Aln());
					if (n[i] && a[i][0].score >= o.T) { mate_hdr[read][0].has = true; add_req(reqs, read, 0, 2, 0, 0); }
				}
				int ef = 1;
				if (n[0] && n[1] && a[0][0].score >= o.T && a[1][0].score >= o.T && a[0][0].rid == a[1][0].rid) {
					i64 dist;
					const int d = infer_dir(idx->l_pac, a[0][0].rb, a[1][0].rb, &dist);
					if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) ef |= 2;
				}
				pi.extra_flag = ef;
				plan_reg2sam(reqs, 2 * p, 0x41 | ef);
				plan_reg2sam(reqs, 2 * p + 1, 0x81 | ef);
			}
		}
	}
	};
	t0 = now();
	{
		const int units = paired ? n_reads / 2 : n_reads;
		std::vector<std::thread> th;
		for (int t = 0; t < n_thr; ++t) {
			const int lo = (int)((long long)units * t / n_thr), hi = (int)((long long)units * (t + 1) / n_thr);
			if (n_thr == 1) plan_range(0, lo, hi); else th.emplace_back(plan_range, t, lo, hi);
		}
		for (size_t t = 0; t < th.size(); ++t) th[t].join();
		for (int t = 0; t < n_thr; ++t) reqs.insert(reqs.end(), reqs_t[t].begin(), reqs_t[t].end()); // thread order = read order
	}
	t_plan = now() - t0;
	const int nt = (int)reqs.size();
	std::vector<AlnOut> outs(nt);
	std::vector<u32> cigs((size_t)nt * CIG_CAP + 1);
	std::vector<char> mds((size_t)nt * MD_CAP + 1);
	if (nt) {
		std::vector<CigTask> tasks(nt);
		for (int t = 0; t < nt; ++t) { tasks[t].reg = reqs[t].reg; tasks[t].read = reqs[t].read; tasks[t].pad = 0; }
		if ((rc = be.cigar(tasks, outs, cigs, mds))) return rc;
	}
	// 7. distribute results: output lines, mate headers, XA strings
	t0 = now();
	std::vector<std::vector<std::string> > xa(n_reads);
	{
		// requests are in read order and thread t planned the reads of units [lo_t, hi_t): its slice of `reqs` touches only those
		// reads' lines, mate headers and XA strings, so the slices can be distributed concurrently
		std::vector<size_t> req_lo(n_thr + 1, 0);
		for (int t = 0; t < n_thr; ++t) req_lo[t + 1] = req_lo[t] + reqs_t[t].size();
		const int units = paired ? n_reads / 2 : n_reads, per_unit = paired ? 2 : 1;
		std::vector<int> bad(n_thr, 0);
		auto dist_range = [&](int tid) {
			const int rlo = (int)((long long)units * tid / n_thr) * per_unit, rhi = (int)((long long)units * (tid + 1) / n_thr) * per_unit;
			for (int r = rlo; r < rhi; ++r) xa[r].resize(na[r]);
			for (size_t t = req_lo[tid]; t < req_lo[tid + 1]; ++t) {
				const CigReq &q = reqs[t];
				const AlnOut &ao = outs[t];
				if (ao.n_cigar < 0 || ao.n_cigar > CIG_CAP - 2 || ao.md_len >= MD_CAP) { bad[tid] = 1; return; }
				if (q.kind == 1) {
					std::string &s = xa[q.read][q.xa_owner];
					s += idx->names[ao.rid]; s += ','; s += "+-"[ao.is_rev]; put_num(s, ao.pos + 1); s += ',';
					for (int k = 0; k < ao.n_cigar; ++k) { put_num(s, cigs[(size_t)t * CIG_CAP + k] >> 4); s += "MIDSHN"[cigs[(size_t)t * CIG_CAP + k] & 0xf]; }
					s += ','; put_num(s, ao.NM); s += ';';
					continue;
				}
				Aln &dst = q.kind == 0 ? plan[q.read].lines[q.line] : mate_hdr[q.read][0];
				dst.pos = ao.pos; dst.rid = ao.rid; dst.is_rev = ao.is_rev; dst.NM = ao.NM; dst.score = ao.score; dst.sub = ao.sub;
				dst.flag |= ao.flag;
				dst.cigar.assign(cigs.begin() + (size_t)t * CIG_CAP, cigs.begin() + (size_t)t * CIG_CAP + ao.n_cigar);
				dst.md.assign(mds.data() + (size_t)t * MD_CAP, ao.md_len);
				if (q.kind == 0) plan[q.read].xa_for.resize(1);
			}
			// XA goes to the line built from region xa_owner (all of a read's XA requests precede this pass over its lines)
			for (size_t t = req_lo[tid]; t < req_lo[tid + 1]; ++t) if (reqs[t].kind == 0) plan[reqs[t].read].lines[reqs[t].line].xa = xa[reqs[t].read][reqs[t].reg_idx];
		};
		std::vector<std::thread> th;
		for (int t = 0; t < n_thr; ++t) { if (n_thr == 1) dist_range(0); else th.emplace_back(dist_range, t); }
		for (size_t t = 0; t < th.size(); ++t) th[t].join();
		for (int t = 0; t < n_thr; ++t) if (bad[t]) { err = "alignment with too many CIGAR operations / MD characters or a traceback matrix beyond the per-thread capacity"; return SSQ_ECAP; }
	}
	t_dist = now() - t0;
	// 8. SAM text in input order
	t0 = now();
	sam.clear();
	std::vector<std::string> part(n_thr);
	std::vector<size_t> rel_off(line_off ? n_reads + 1 : 0, 0);
	auto fmt_range = [&](int tid, int lo, int hi) {
		std::string &sam = part[tid];
		sam.reserve((size_t)(hi - lo) * 400);
		for (int r = lo; r < hi; ++r) {
			if (line_off) rel_off[r] = sam.size();
			const char *sq = (const char*)codes.data() + off[r];
			const int l_seq = (int)(off[r + 1] - off[r]);
			const char *cm = comments ? comments[r] : 0;
			const Aln *mate = 0;
			if (paired) {
				const int m = r ^ 1;
				const PairInfo &pi = pinfo[r >> 1];
				if (pi.mode == 0) mate = &plan[m].lines[0];
				else mate = &mate_hdr[m][0]; // has==false -> unmapped header (rid -1)
			}
			if (plan[r].lines.empty()) { // unaligned record
				std::vector<Aln> one(1);
				one[0].flag = 0x4 | plan[r].extra_flag;
				aln2sam(idx, sam, names[r], sq, l_seq, quals ? quals[r] : 0, one, 0, mate, rg_id, cm);
			} else {
				for (size_t k = 0; k < plan[r].lines.size(); ++k) aln2sam(idx, sam, names[r], sq, l_seq, quals ? quals[r] : 0, plan[r].lines, (int)k, mate, rg_id, cm);
			}
		}
	};
	{
		std::vector<std::thread> th;
		std::vector<int> lo_of(n_thr + 1, 0);
		for (int t = 0; t <= n_thr; ++t) lo_of[t] = (int)((long long)n_reads * t / n_thr);
		for (int t = 0; t < n_thr; ++t) { if (n_thr == 1) fmt_range(0, lo_of[t], lo_of[t + 1]); else th.emplace_back(fmt_range, t, lo_of[t], lo_of[t + 1]); }
		for (size_t t = 0; t < th.size(); ++t) th[t].join();
		size_t total = 0;
		for (int t = 0; t < n_thr; ++t) total += part[t].size();
		sam.reserve(total + 1);
		if (line_off) line_off->assign(n_reads + 1, 0);
		for (int t = 0; t < n_thr; ++t) {
			if (line_off) for (int r = lo_of[t]; r < lo_of[t + 1]; ++r) (*line_off)[r] = sam.size() + rel_off[r];
			sam += part[t];
		}
	}
	if (line_off) (*line_off)[n_reads] = sam.size();
	t_fmt = now() - t0;
	if (timing) fprintf(stderr, "[ssq_mem host] %d reads, %d threads: plan (primary/pair/MAPQ) %.3f s | distribute results %.3f s | SAM text %.3f s\n", n_reads, n_thr, t_plan, t_dist, t_fmt);
	return SSQ_OK;
}
