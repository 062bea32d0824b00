/* TEST-ONLY stand-in for libssq.so's ssq_dupset_* entry points so that the text logic of the `samblaster` shim
 * (speedseq_b200/cli/samblaster_main.c) can be fuzzed against the oracle on a box without a GPU.  Never shipped, never linked
 * into the product: tests/test_samblaster_shim_cpu.py builds the shim against this file into a temporary directory. */
#include <stdlib.h>
#include <string.h>
#include "ssq.h"
struct ssq_dupset { ssq_dupsig_t *v; size_t n, m; };
const char *ssq_last_error(void) { return "stub"; }
int ssq_dupset_create(int device, ssq_dupset_t **out) { (void)device; *out = (ssq_dupset_t*)calloc(1, sizeof(**out)); return 0; }
void ssq_dupset_free(ssq_dupset_t *s) { if (s) { free(s->v); free(s); } }
uint64_t ssq_dupset_size(const ssq_dupset_t *s) { return s->n; }
int ssq_dupset_mark(ssq_dupset_t *s, uint64_t n, const ssq_dupsig_t *sig, uint8_t *is_dup)
{
	for (uint64_t i = 0; i < n; ++i) {
		size_t k;
		is_dup[i] = 0;
		if (!sig[i].valid) continue;
		for (k = 0; k < s->n; ++k)
			if (s->v[k].pos1 == sig[i].pos1 && s->v[k].pos2 == sig[i].pos2 && s->v[k].strand1 == sig[i].strand1 && s->v[k].strand2 == sig[i].strand2) break;
		if (k < s->n) { is_dup[i] = 1; continue; }
		if (s->n == s->m) { s->m = s->m ? s->m * 2 : 1024; s->v = (ssq_dupsig_t*)realloc(s->v, s->m * sizeof(ssq_dupsig_t)); }
		s->v[s->n++] = sig[i];
	}
	return 0;
}
