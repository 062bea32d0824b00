// ssq_host.h — host-side definitions shared by the translation units of libssq.so.
#pragma once
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
