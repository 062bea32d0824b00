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
