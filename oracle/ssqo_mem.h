/* ssqo_mem.h — ORACLE (test infrastructure): internals shared by ssqo_mem.c and ssqo_pair.c. */
#ifndef SSQO_MEM_H
#define SSQO_MEM_H
#include "ssqo.h"

typedef struct { size_t l, m; char *s; } ssqo_sb_t;

typedef struct { /* one SAM-ready alignment */
	int64_t pos;
	int rid, flag;
	uint32_t is_rev:1, is_alt:1, mapq:8, NM:22;
	int n_cigar;
	uint32_t *cigar;
	char *md, *XA;
	int score, sub, alt_sc;
} ssqo_aln_t;

void ssqo_sort_u64(size_t n, uint64_t *a);
uint8_t *ssqo_fetch_seq(const ssqo_bns_t *bns, const uint8_t *pac, int64_t *beg, int64_t mid, int64_t *end, int *rid);
uint32_t *ssqo_gen_cigar2(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *pac,
                          int l_query, uint8_t *query, int64_t rb, int64_t re, int *score, int *n_cigar, int *NM, char **md);
ssqo_aln_t ssqo_reg2aln(const ssqo_opt_t *opt, const ssqo_idx_t *idx, int l_query, const char *query, const ssqo_alnreg_t *ar);
void ssqo_aln_free(ssqo_aln_t *a);
void ssqo_aln2sam(const ssqo_opt_t *opt, const ssqo_bns_t *bns, ssqo_sb_t *str, const ssqo_read_t *s, int n, const ssqo_aln_t *list, int which, const ssqo_aln_t *m, const char *rg_id);
char **ssqo_gen_alt(const ssqo_opt_t *opt, const ssqo_idx_t *idx, const ssqo_alnreg_v *a, int l_query, const char *query);
void ssqo_reg2sam(const ssqo_opt_t *opt, const ssqo_idx_t *idx, ssqo_read_t *s, ssqo_alnreg_v *a, int extra_flag, const ssqo_aln_t *m, const char *rg_id);
#endif
