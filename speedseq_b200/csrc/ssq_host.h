// ssq_host.h — host-side definitions shared by the translation units of libssq.so.
#pragma once
#include <cuda_runtime.h>
#include "ssq_dev.cuh"

struct ssq_index {
	int device;
	DevIndex dev;          // device pointers + scalars, passed by value to kernels
	size_t dev_bytes;
	// host copies needed by the CLI (SAM header, rid names)
	int n_seqs;
	char **names;
	i64 *ann_off;
	i32 *ann_len;
};

void ssq_set_error(const char *fmt, ...);
int ssq_use_device(int device); // cudaSetDevice + architecture check; SSQ_ENOGPU if unusable

struct DBuf { // growable device buffer
	void *p; size_t cap;
	DBuf() : p(0), cap(0) {}
	~DBuf() { release(); }
	int need(size_t bytes) {
		if (bytes <= cap) return 0;
		if (p) cudaFree(p);
		size_t want = bytes + bytes / 4 + 256;
		if (cudaMalloc(&p, want) != cudaSuccess) { p = 0; cap = 0; ssq_set_error("cudaMalloc(%zu) failed", want); return SSQ_ENOMEM; }
		cap = want; return 0;
	}
	void release() { if (p) cudaFree(p); p = 0; cap = 0; }
	template <class T> T *as() { return (T*)p; }
};
