// ssq_mem.cu — `bwa mem` for a batch of reads: GPU stages + host orchestration -> SAM text.
//
// Reference call site: `$BWA mem -t T [-p] -R RG REF FQ` at /root/reference/bin/speedseq:438,468 (options all default).
// Upstream routines replaced (not vendored): mem_process_seqs, mem_pestat, mem_sam_pe, mem_matesw, mem_pair,
// mem_mark_primary_se, mem_approx_mapq_se, mem_reg2aln, mem_gen_alt, mem_reg2sam, mem_aln2sam.
//
// Division of labour: everything that touches bases or DP cells runs on the GPU —
//   ssq_batch_run            seeding, SA, chaining, extension            (ssq_kernels.cu)
//   k_dedup                  sort / de-duplicate / patch regions          (thread per read, banded global DP in-thread)
//   k_matesw                 mate rescue, striped-order local SW          (thread per pair)
//   k_cigar                  CIGAR / NM / MD by banded global DP+traceback (thread per reported alignment)
// and the host keeps what is O(#hits) bookkeeping on doubles and text: insert-size statistics (needs IEEE libm results
// identical to the reference: erfc/log), primary marking, pairing, MAPQ, and SAM formatting.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <chrono>
#include "ssq_mem_host.h"
#include "ssq_host.h"
#include "ssq_batch.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ssq_set_error("%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return SSQ_ECUDA; } } while (0)

// ================================================================================ kernels ====
#define REF_CAP 12288    // longest reference window a thread materialises (mate rescue: max_ins + 2 reads)

struct ThreadScratch { // carved from a per-thread slab
	uint8_t *qbuf, *rbuf, *z, *seq, *ref; i32 *h, *e, *H0, *H1, *E, *Hmax; u64 *b;
};
#define QMAX 256
#define Z_BYTES (QMAX * 768)
// big slab (mate rescue, CIGAR): the traceback matrix (k_cigar) and the rescue window + sub-optimal row list (k_matesw) are
// never live together, so they share one area; small slab (k_dedup): score-only global DP needs two DP rows and two sequences
#define BIG_AREA (Z_BYTES > REF_CAP * 9 ? Z_BYTES : REF_CAP * 9)
#define SLAB_FIXED (6 * (QMAX + 16) * 4 + QMAX + 2048 + QMAX + 64)
#define SLAB_BYTES (SLAB_FIXED + BIG_AREA)
#define SLAB_SMALL_BYTES (2 * (QMAX + 16) * 4 + QMAX + 2048 + 64)

__device__ __forceinline__ void carve(uint8_t *slab, ThreadScratch &t)
{
	uint8_t *p = slab;
	t.h = (i32*)p; p += (QMAX + 16) * 4; t.e = (i32*)p; p += (QMAX + 16) * 4;
	t.H0 = (i32*)p; p += (QMAX + 16) * 4; t.H1 = (i32*)p; p += (QMAX + 16) * 4; t.E = (i32*)p; p += (QMAX + 16) * 4; t.Hmax = (i32*)p; p += (QMAX + 16) * 4;
	t.qbuf = p; p += QMAX; t.rbuf = p; p += 2048; t.seq = p; p += QMAX + 64;
	t.z = p;                                   // k_cigar's view of the shared area
	t.b = (u64*)p; t.ref = p + (size_t)REF_CAP * 8; // k_matesw's view
}
__device__ __forceinline__ void carve_small(uint8_t *slab, ThreadScratch &t)
{
	uint8_t *p = slab;
	t.h = (i32*)p; p += (QMAX + 16) * 4; t.e = (i32*)p; p += (QMAX + 16) * 4;
	t.qbuf = p; p += QMAX; t.rbuf = p;
	t.H0 = t.H1 = t.E = t.Hmax = 0; t.seq = t.ref = t.z = 0; t.b = 0;
}
__device__ __forceinline__ void scratch_views(const ThreadScratch &t, AlnScratch &A, MateScratch &M)
{
	A.qbuf = t.qbuf; A.rbuf = t.rbuf; A.rcap = 2048; A.g.h = t.h; A.g.e = t.e; A.g.z = t.z; A.g.zcap = t.z ? Z_BYTES : 0;
	M.seq = t.seq; M.ref = t.ref; M.ref_cap = REF_CAP; M.L.H0 = t.H0; M.L.H1 = t.H1; M.L.E = t.E; M.L.Hmax = t.Hmax; M.L.b = t.b; M.L.b_cap = REF_CAP; M.A = A;
}

// stage 1: per read, RegCand list -> AlnReg list (sorted, de-duplicated, patched) at areg[areg_off[r] ..), count in n_areg[r]
__global__ void __launch_bounds__(128) k_dedup(DevIndex ix, ssq_opts_t opt, int n_reads, const uint8_t *__restrict__ seq, const u64 *__restrict__ read_off,
                                               const u64 *__restrict__ task_off, const u32 *__restrict__ n_regs, const RegCand *__restrict__ regs,
                                               const u64 *__restrict__ areg_off, AlnReg *areg, u32 *n_areg, uint8_t *slabs, int *work)
{
	ThreadScratch ts; AlnScratch A; MateScratch M;
	carve_small(slabs + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * SLAB_SMALL_BYTES, ts);
	scratch_views(ts, A, M);
	for (;;) {
		const int r = atomicAdd(work, 1);
		if (r >= n_reads) break;
		const int n0 = (int)n_regs[r];
		AlnReg *a = areg + areg_off[r];
		for (int i = 0; i < n0; ++i) reg_from_cand(regs[task_off[r] + i], a[i]);
		n_areg[r] = (u32)sort_dedup_patch(ix, opt, seq + read_off[r], n0, a, A);
	}
}

// stage 2: per pair, mate rescue from the near-best hits of either end (mem_sam_pe's first block)
__global__ void __launch_bounds__(128) k_matesw(DevIndex ix, ssq_opts_t opt, int n_pairs, const uint8_t *__restrict__ seq, const u64 *__restrict__ read_off,
                                                const u64 *__restrict__ areg_off, AlnReg *areg, u32 *n_areg, const PeStat *__restrict__ pes_, uint8_t *slabs,
                                                AlnReg *bbuf /* per thread 2 x 64 */, int *work, int *err)
{
	ThreadScratch ts; AlnScratch A; MateScratch M;
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	carve(slabs + tid * SLAB_BYTES, ts);
	scratch_views(ts, A, M);
	PeStat pes[4];
	for (int i = 0; i < 4; ++i) pes[i] = pes_[i];
	AlnReg *b[2] = {bbuf + tid * 128, bbuf + tid * 128 + 64};
	for (;;) {
		const int p = atomicAdd(work, 1);
		if (p >= n_pairs) break;
		int nb[2] = {0, 0}, na[2];
		AlnReg *a[2];
		for (int i = 0; i < 2; ++i) {
			a[i] = areg + areg_off[2 * p + i]; na[i] = (int)n_areg[2 * p + i];
			for (int j = 0; j < na[i]; ++j)
				if (a[i][j].score >= a[i][0].score - opt.pen_unpaired && nb[i] < 64) b[i][nb[i]++] = a[i][j]; // only the first max_matesw (50) are used
		}
		for (int i = 0; i < 2; ++i) {
			const int cap = (int)(areg_off[2 * p + !i + 1] - areg_off[2 * p + !i]);
			for (int j = 0; j < nb[i] && j < opt.max_matesw; ++j) {
				const int before = na[!i];
				mate_rescue(ix, opt, pes, b[i][j], (int)(read_off[2 * p + !i + 1] - read_off[2 * p + !i]), seq + read_off[2 * p + !i], a[!i], &na[!i], cap, M);
				if (na[!i] >= cap && before < cap) atomicMax(err, 1);
			}
		}
		n_areg[2 * p] = (u32)na[0]; n_areg[2 * p + 1] = (u32)na[1];
	}
}

__global__ void __launch_bounds__(128) k_cigar(DevIndex ix, ssq_opts_t opt, int n_tasks, const CigTask *__restrict__ tasks, const uint8_t *__restrict__ seq,
                                               const u64 *__restrict__ read_off, AlnOut *out, u32 *cig_pool, char *md_pool, uint8_t *slabs, int *work)
{
	ThreadScratch ts; AlnScratch A; MateScratch M;
	carve(slabs + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * SLAB_BYTES, ts);
	scratch_views(ts, A, M);
	for (;;) {
		const int t = atomicAdd(work, 1);
		if (t >= n_tasks) break;
		const CigTask k = tasks[t];
		AlnOut a;
		reg2aln(ix, opt, (int)(read_off[k.read + 1] - read_off[k.read]), seq + read_off[k.read], k.reg, A, a, cig_pool + (size_t)t * CIG_CAP, CIG_CAP, md_pool + (size_t)t * MD_CAP, MD_CAP);
		out[t] = a;
	}
}

// capacity of a read's region list: its own regions plus at most 4 rescued ones per mate hit that may trigger a rescue
__global__ void k_areg_cap(int n, int paired, int max_matesw, const u32 *__restrict__ n_regs, u64 *cap)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	u32 m = paired ? n_regs[i ^ 1] : 0;
	if (m > (u32)max_matesw) m = (u32)max_matesw;
	cap[i] = (u64)n_regs[i] + 4ull * m + (paired ? 4 : 0);
}

// =========================================================================== CUDA backend ====
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct CudaBackend {
	double t_align, t_dedup, t_rescue, t_cigar, t_alloc;
	const ssq_index *idx; ssq_opts_t o;
	ssq_batch_t *b; cudaStream_t st; BatchView bv;
	std::vector<void*> dev;
	uint8_t *d_slabs, *d_slabs_small; u64 *d_cap, *d_aoff; u32 *d_na; AlnReg *d_areg; int *d_work; int n_reads, slab_threads; u64 total_cap;
	std::vector<u64> aoff; std::vector<u32> na; std::vector<AlnReg> areg;
	CudaBackend(const ssq_index *i, const ssq_opts_t &o_) : idx(i), o(o_), b(0), d_slabs(0), d_slabs_small(0), d_cap(0), d_aoff(0), d_na(0), d_areg(0), d_work(0), n_reads(0), slab_threads(0), total_cap(0) { t_align = t_dedup = t_rescue = t_cigar = t_alloc = 0; }
	~CudaBackend() { for (size_t i = 0; i < dev.size(); ++i) cudaFree(dev[i]); if (b) ssq_batch_free(b); }
#define DMALLOC(p, bytes) do { CK(cudaMalloc((void**)&(p), (bytes))); dev.push_back((void*)(p)); } while (0)
	int fetch()
	{
		CK(cudaMemcpyAsync(aoff.data(), d_aoff, (size_t)(n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(na.data(), d_na, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(areg.data(), d_areg, total_cap * sizeof(AlnReg), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		return 0;
	}
	int align(int n, const uint8_t *codes, const u64 *off, int paired, int max_matesw)
	{
		int rc;
		n_reads = n;
		double t0 = now_s();
		if ((rc = ssq_batch_create(idx, &o, n, codes, off, &b))) return rc;
		if ((rc = ssq_batch_run(b))) return rc;
		ssq_batch_sync(b); t_align = now_s() - t0; t0 = now_s();
		st = (cudaStream_t)ssq_batch_stream(b);
		bv = ssq_batch_view(b);
		slab_threads = bv.n_sm * 128;            // big slabs: one block per SM
		const int small_threads = bv.n_sm * 8 * 128; // small slabs: the dedup kernel runs wide
		void *d_tmp = 0; size_t tmp_bytes = 0;
		DMALLOC(d_slabs, (size_t)slab_threads * SLAB_BYTES);
		DMALLOC(d_slabs_small, (size_t)small_threads * SLAB_SMALL_BYTES);
		t_alloc = now_s() - t0; t0 = now_s();
		DMALLOC(d_cap, (size_t)(n + 2) * 8); DMALLOC(d_aoff, (size_t)(n + 2) * 8); DMALLOC(d_na, (size_t)(n + 1) * 4); DMALLOC(d_work, 64);
		CK(cudaMemsetAsync(d_aoff, 0, (size_t)(n + 2) * 8, st));
		if (n) {
			k_areg_cap<<<(n + 255) / 256, 256, 0, st>>>(n, paired, max_matesw, bv.n_regs, d_cap);
			cub::DeviceScan::ExclusiveSum(0, tmp_bytes, d_cap, d_aoff, n + 1, st);
			DMALLOC(d_tmp, tmp_bytes);
			CK(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cap, d_aoff, n + 1, st));
			CK(cudaMemcpyAsync(&total_cap, d_aoff + n, 8, cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
		}
		DMALLOC(d_areg, (total_cap + 1) * sizeof(AlnReg));
		CK(cudaMemsetAsync(d_work, 0, 64, st));
		if (n) k_dedup<<<bv.n_sm * 8, 128, 0, st>>>(bv.ix, o, n, bv.seq, bv.read_off, bv.task_off, bv.n_regs, bv.regs, d_aoff, d_areg, d_na, d_slabs_small, d_work);
		CK(cudaGetLastError());
		aoff.assign(n + 1, 0); na.assign(n + 1, 0); areg.resize(total_cap + 1);
		rc = fetch(); t_dedup = now_s() - t0;
		return rc;
	}
	int rescue(const PeStat pes[4])
	{
		double t0 = now_s();
		PeStat *d_pes = 0; AlnReg *d_bbuf = 0; int *d_err = d_work + 8, h_err = 0;
		DMALLOC(d_pes, 4 * sizeof(PeStat)); DMALLOC(d_bbuf, (size_t)slab_threads * 128 * sizeof(AlnReg));
		CK(cudaMemcpyAsync(d_pes, pes, 4 * sizeof(PeStat), cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(d_work, 0, 64, st));
		k_matesw<<<bv.n_sm, 128, 0, st>>>(bv.ix, o, n_reads / 2, bv.seq, bv.read_off, d_aoff, d_areg, d_na, d_pes, d_slabs, d_bbuf, d_work, d_err);
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		if (h_err) { ssq_set_error("mate rescue overflowed a region list (internal capacity rule violated)"); return SSQ_ECAP; }
		int rc = fetch(); t_rescue = now_s() - t0;
		return rc;
	}
	int cigar(const std::vector<CigTask> &tasks, std::vector<AlnOut> &outs, std::vector<u32> &cigs, std::vector<char> &mds)
	{
		const int nt = (int)tasks.size();
		double t0 = now_s();
		CigTask *d_tasks = 0; AlnOut *d_out = 0; u32 *d_cig = 0; char *d_md = 0;
		DMALLOC(d_tasks, (size_t)nt * sizeof(CigTask)); DMALLOC(d_out, (size_t)nt * sizeof(AlnOut)); DMALLOC(d_cig, (size_t)nt * CIG_CAP * 4); DMALLOC(d_md, (size_t)nt * MD_CAP);
		CK(cudaMemcpyAsync(d_tasks, tasks.data(), (size_t)nt * sizeof(CigTask), cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(d_work, 0, 64, st));
		k_cigar<<<bv.n_sm, 128, 0, st>>>(bv.ix, o, nt, d_tasks, bv.seq, bv.read_off, d_out, d_cig, d_md, d_slabs, d_work);
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(outs.data(), d_out, (size_t)nt * sizeof(AlnOut), cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(cigs.data(), d_cig, (size_t)nt * CIG_CAP * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(mds.data(), d_md, (size_t)nt * MD_CAP, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		t_cigar = now_s() - t0;
		return 0;
	}
};

extern "C" int ssq_mem_batch_sam(const ssq_index_t *idx, const ssq_opts_t *opt_, int n_reads, const char *const *names, const char *const *seqs, const char *const *quals,
                                 const char *const *comments, int64_t n_processed, int paired, const ssq_pestat_t *pes0, const char *rg_id, int verbose, char **sam_out, size_t *sam_len, size_t *read_sam_off)
{
	if (!idx || !opt_ || n_reads < 0 || !sam_out) return SSQ_EINVAL;
	if (paired && (n_reads & 1)) { ssq_set_error("paired batch with an odd number of reads"); return SSQ_EINVAL; }
	int rc = ssq_use_device(idx->device);
	if (rc) return rc;
	std::vector<u64> off(n_reads + 1, 0);
	for (int i = 0; i < n_reads; ++i) off[i + 1] = off[i] + strlen(seqs[i]);
	std::vector<uint8_t> codes(off[n_reads] + 1);
	{
		uint8_t lut[256];
		memset(lut, 4, sizeof lut);
		lut['A'] = lut['a'] = 0; lut['C'] = lut['c'] = 1; lut['G'] = lut['g'] = 2; lut['T'] = lut['t'] = 3;
		for (int i = 0; i < n_reads; ++i) {
			const unsigned char *s = (const unsigned char*)seqs[i];
			uint8_t *d = codes.data() + off[i];
			const size_t l = (size_t)(off[i + 1] - off[i]);
			for (size_t k = 0; k < l; ++k) d[k] = lut[s[k]];
		}
	}
	HostIndexInfo hi; hi.l_pac = idx->dev.l_pac; hi.n_seqs = idx->n_seqs; hi.names = idx->names; hi.ann_off = idx->ann_off;
	PeStat pes[4];
	if (pes0) for (int d = 0; d < 4; ++d) { pes[d].low = pes0[d].low; pes[d].high = pes0[d].high; pes[d].failed = pes0[d].failed; pes[d].pad = 0; pes[d].avg = pes0[d].avg; pes[d].std = pes0[d].std; }
	CudaBackend be(idx, *opt_);
	std::string sam, err;
	std::vector<size_t> loff;
	const double t_all0 = now_s();
	rc = mem_batch_sam(be, *opt_, &hi, n_reads, names, codes.data(), off.data(), quals, comments, n_processed, paired, pes0 ? pes : 0, rg_id, verbose ? stderr : 0, sam, err, read_sam_off ? &loff : 0);
	if (getenv("SSQ_MEM_TIMING")) fprintf(stderr, "[ssq_mem] %d reads: total %.3f s | seed..extend %.3f | slab alloc %.3f | dedup+fetch %.3f | rescue+fetch %.3f | cigar %.3f | host rest %.3f\n", n_reads, now_s() - t_all0, be.t_align, be.t_alloc, be.t_dedup, be.t_rescue, be.t_cigar, now_s() - t_all0 - be.t_align - be.t_alloc - be.t_dedup - be.t_rescue - be.t_cigar);
	if (rc) { if (!err.empty()) ssq_set_error("%s", err.c_str()); return rc; }
	*sam_out = (char*)malloc(sam.size() + 1);
	if (!*sam_out) return SSQ_ENOMEM;
	memcpy(*sam_out, sam.c_str(), sam.size() + 1);
	if (sam_len) *sam_len = sam.size();
	if (read_sam_off) memcpy(read_sam_off, loff.data(), (size_t)(n_reads + 1) * sizeof(size_t));
	return SSQ_OK;
}

extern "C" void ssq_free(void *p) { free(p); }
