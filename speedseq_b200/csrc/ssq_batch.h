// ssq_batch.h — read-only view of a batch object's device buffers for the later stages (ssq_mem.cu)
#pragma once
#include "ssq_dev.cuh"
struct BatchView {
	DevIndex ix;
	int n_reads, n_sm;
	const uint8_t *seq; const u64 *read_off;
	const u64 *task_off; const u32 *n_regs; const RegCand *regs; // stage-0 regions of read r: regs[task_off[r] .. +n_regs[r])
};
BatchView ssq_batch_view(ssq_batch_t *b);
// sizes the read buffers of a batch object for n_reads reads / total_bases bases (longest read max_len) WITHOUT copying: the caller
// fills d_seq (one code per base) and d_off (n_reads + 1 offsets) on the batch's stream (ssq_pipe.cu converts ASCII on the device)
int ssq_batch_reserve(ssq_batch_t *b, int n_reads, u64 total_bases, int max_len, uint8_t **d_seq, u64 **d_off);
