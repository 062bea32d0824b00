/*
 * bwa — drop-in for the `$BWA` entry of speedseq.config (/root/reference/bin/speedseq.config:13), a thin C shim over
 * libssq.so.  Honours the argv/stdio contract of the reference's call sites:
 *     $BWA index REF                                              /root/reference/bin/speedseq:389
 *     $BWA mem -t T [-p] [-C] [-I f[,f[,i[,i]]]] -R '@RG\tID:..' REF FQ1 [FQ2]   speedseq:438,468,1961
 * stdout = SAM (header, then records name-grouped in input order).  All base-level work happens on the GPU inside
 * libssq (ssq_index_build, ssq_mem_batch_sam); this file only parses argv, tokenises FASTQ, forms batches the way the
 * reference's `bwa mem` does (bases >= 10 M x T and an even read count) and writes text.
 * FASTQ tokenisation follows the reference's in-tree parser /root/reference/src/samtools-1.3.1/htslib-1.3.1/htslib/kseq.h:189-231.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>
#include "ssq.h"

#define SHIM_VERSION "0.7.12-r1039" /* the bwa release whose behaviour libssq reproduces */

/* ---------------------------------------------------------------- FASTQ/FASTA reader ---- */
typedef struct { gzFile fp; unsigned char *buf; int beg, end, eof, last; } fq_t;
typedef struct { char *s; size_t l, m; } str_t;
typedef struct { str_t name, comment, seq, qual; } rec_t;

static fq_t *fq_open(const char *fn)
{
	gzFile fp = strcmp(fn, "-") ? gzopen(fn, "r") : gzdopen(0, "r");
	fq_t *f;
	if (!fp) return 0;
	f = (fq_t*)calloc(1, sizeof(fq_t));
	f->fp = fp; f->buf = (unsigned char*)malloc(1 << 16);
	return f;
}
static inline int fq_getc(fq_t *f)
{
	if (f->beg >= f->end) {
		if (f->eof) return -1;
		f->beg = 0; f->end = gzread(f->fp, f->buf, 1 << 16);
		if (f->end <= 0) { f->eof = 1; f->end = 0; return -1; }
	}
	return f->buf[f->beg++];
}
static inline void s_push(str_t *s, int c)
{
	if (s->l + 2 > s->m) { s->m = s->m ? s->m * 2 : 128; s->s = (char*)realloc(s->s, s->m); }
	s->s[s->l++] = (char)c; s->s[s->l] = 0;
}
static inline void s_clear(str_t *s) { s->l = 0; if (!s->s) { s->m = 128; s->s = (char*)malloc(s->m); } s->s[0] = 0; }
static int fq_line(fq_t *f, str_t *s) /* appends the rest of the current line, returns -1 at EOF with nothing read */
{
	int c, got = 0;
	while ((c = fq_getc(f)) >= 0) { got = 1; if (c == '\n') break; s_push(s, c); }
	if (s->l && s->s[s->l - 1] == '\r') s->s[--s->l] = 0;
	return got ? 0 : -1;
}
/* >= 0: sequence length; -1: end of file; -2: truncated quality */
static int fq_read(fq_t *f, rec_t *r)
{
	int c;
	if (f->last == 0) {
		while ((c = fq_getc(f)) >= 0 && c != '>' && c != '@');
		if (c < 0) return -1;
		f->last = c;
	}
	s_clear(&r->name); s_clear(&r->comment); s_clear(&r->seq); s_clear(&r->qual);
	while ((c = fq_getc(f)) >= 0 && !isspace(c)) s_push(&r->name, c);
	if (c < 0 && r->name.l == 0) return -1;
	if (c >= 0 && c != '\n') fq_line(f, &r->comment);
	while ((c = fq_getc(f)) >= 0 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		s_push(&r->seq, c);
		fq_line(f, &r->seq);
	}
	if (c == '>' || c == '@') f->last = c;
	if (c != '+') { if (c < 0) f->last = 0; return (int)r->seq.l; }
	while ((c = fq_getc(f)) >= 0 && c != '\n');
	if (c < 0) return -2;
	while (r->qual.l < r->seq.l && fq_line(f, &r->qual) == 0);
	f->last = 0;
	return r->seq.l == r->qual.l ? (int)r->seq.l : -2;
}

/* --------------------------------------------------------------------------- batches ---- */
typedef struct { char *name, *comment, *seq, *qual; int id; } read_t;
typedef struct { read_t *a; int n, m; } reads_t;

static void trim_readno(str_t *s) { if (s->l > 2 && s->s[s->l - 2] == '/' && isdigit((unsigned char)s->s[s->l - 1])) { s->l -= 2; s->s[s->l] = 0; } }

static void reads_push(reads_t *v, const rec_t *r, int keep_comment)
{
	read_t *x;
	if (v->n == v->m) { v->m = v->m ? v->m * 2 : 1024; v->a = (read_t*)realloc(v->a, sizeof(read_t) * v->m); }
	x = &v->a[v->n];
	x->name = strdup(r->name.s); x->seq = strdup(r->seq.s);
	x->qual = r->qual.l ? strdup(r->qual.s) : 0;
	x->comment = keep_comment && r->comment.l ? strdup(r->comment.s) : 0;
	x->id = v->n++;
}

/* one batch: until the base count reaches chunk and the read count is even */
static long read_batch(long chunk, fq_t *f1, fq_t *f2, rec_t *r1, rec_t *r2, reads_t *v, int keep_comment)
{
	long size = 0;
	v->n = 0;
	while (fq_read(f1, r1) >= 0) {
		if (f2 && fq_read(f2, r2) < 0) { fprintf(stderr, "[W::bseq_read] the 2nd file has fewer sequences.\n"); break; }
		trim_readno(&r1->name); reads_push(v, r1, keep_comment); size += (long)r1->seq.l;
		if (f2) { trim_readno(&r2->name); reads_push(v, r2, keep_comment); size += (long)r2->seq.l; }
		if (size >= chunk && (v->n & 1) == 0) break;
	}
	if (size == 0 && f2 && fq_read(f2, r2) >= 0) fprintf(stderr, "[W::bseq_read] the 1st file has fewer sequences.\n");
	return size;
}

static void die(const char *what, int rc) { fprintf(stderr, "[E::bwa] %s failed (%d): %s\n", what, rc, ssq_last_error()); exit(1); }

/* runs one homogeneous (all single-end or all paired) sub-batch and scatters each read's SAM lines to out[id] */
static void run_sub(const ssq_index_t *idx, const ssq_opts_t *opt, read_t **sub, int n, long long n_processed, int paired, const ssq_pestat_t *pes0,
                    const char *rg_id, char **out)
{
	const char **names = (const char**)malloc(sizeof(char*) * n), **seqs = (const char**)malloc(sizeof(char*) * n);
	const char **quals = (const char**)malloc(sizeof(char*) * n), **comments = (const char**)malloc(sizeof(char*) * n);
	size_t *offs = (size_t*)malloc(sizeof(size_t) * (n + 1)), len = 0;
	char *sam = 0;
	int i, rc, any_comment = 0;
	for (i = 0; i < n; ++i) { names[i] = sub[i]->name; seqs[i] = sub[i]->seq; quals[i] = sub[i]->qual; comments[i] = sub[i]->comment; any_comment |= sub[i]->comment != 0; }
	if (paired) for (i = 0; i < n; i += 2) if (strcmp(names[i], names[i + 1]) != 0) { fprintf(stderr, "[E::mem_sam_pe] paired reads have different names: \"%s\", \"%s\"\n", names[i], names[i + 1]); exit(1); }
	rc = ssq_mem_batch_sam(idx, opt, n, names, seqs, quals, any_comment ? comments : 0, n_processed, paired, pes0, rg_id, 1, &sam, &len, offs);
	if (rc) die("ssq_mem_batch_sam", rc);
	for (i = 0; i < n; ++i) {
		const size_t l = offs[i + 1] - offs[i];
		out[sub[i]->id] = (char*)malloc(l + 1);
		memcpy(out[sub[i]->id], sam + offs[i], l); out[sub[i]->id][l] = 0;
	}
	ssq_free(sam);
	free(names); free(seqs); free(quals); free(comments); free(offs);
}

static char *unescape(char *s)
{
	char *p, *q;
	for (p = q = s; *p; ++p) {
		if (*p == '\\') {
			++p;
			if (*p == 't') *q++ = '\t'; else if (*p == 'n') *q++ = '\n'; else if (*p == 'r') *q++ = '\r'; else if (*p == '\\') *q++ = '\\'; else if (*p == 0) break;
		} else *q++ = *p;
	}
	*q = 0;
	return s;
}

static int main_mem(int argc, char **argv, const char *prog)
{
	ssq_opts_t opt;
	ssq_pestat_t pes[4], *pes0 = 0;
	ssq_index_t *idx = 0;
	fq_t *f1, *f2 = 0;
	rec_t r1, r2;
	reads_t v = {0, 0, 0};
	char *rg_line = 0, rg_id[256] = "", *p;
	int c, i, n_threads = 1, smart_pe = 0, paired = 0, keep_comment = 0, device = getenv("SSQ_DEVICE") ? atoi(getenv("SSQ_DEVICE")) : 0, rc;
	long long n_processed = 0;
	const long chunk_size = 10000000;
	memset(&r1, 0, sizeof r1); memset(&r2, 0, sizeof r2); memset(pes, 0, sizeof pes);
	pes[0].failed = pes[1].failed = pes[2].failed = pes[3].failed = 1;
	ssq_opts_default(&opt);
	while ((c = getopt(argc, argv, "t:pR:I:Cv:")) >= 0) {
		if (c == 't') n_threads = atoi(optarg) > 1 ? atoi(optarg) : 1;
		else if (c == 'p') smart_pe = paired = 1;
		else if (c == 'C') keep_comment = 1;
		else if (c == 'v') ;
		else if (c == 'R') {
			if (strstr(optarg, "@RG") != optarg) { fprintf(stderr, "[E::bwa_set_rg] the read group line is not started with @RG\n"); return 1; }
			rg_line = unescape(strdup(optarg));
			if (!(p = strstr(rg_line, "\tID:"))) { fprintf(stderr, "[E::bwa_set_rg] no ID at the read group line\n"); return 1; }
			for (p += 4, i = 0; p[i] && p[i] != '\t' && p[i] != '\n' && i < 255; ++i) rg_id[i] = p[i];
			rg_id[i] = 0;
		} else if (c == 'I') { /* mean[,std[,max[,min]]] for the FR orientation */
			pes0 = pes; pes[1].failed = 0;
			pes[1].avg = strtod(optarg, &p); pes[1].std = pes[1].avg * .1;
			if (*p && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes[1].std = strtod(p + 1, &p);
			pes[1].high = (int)(pes[1].avg + 4. * pes[1].std + .499);
			pes[1].low = (int)(pes[1].avg - 4. * pes[1].std + .499);
			if (pes[1].low < 1) pes[1].low = 1;
			if (*p && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes[1].high = (int)(strtod(p + 1, &p) + .499);
			if (*p && ispunct((unsigned char)*p) && isdigit((unsigned char)p[1])) pes[1].low = (int)(strtod(p + 1, &p) + .499);
		} else return 1;
	}
	opt.n_threads = n_threads;
	if (optind + 1 >= argc || optind + 3 < argc) { fprintf(stderr, "Usage: bwa mem [-t INT] [-p] [-C] [-I FLOAT[,FLOAT[,INT[,INT]]]] [-R STR] <idxbase> <in1.fq> [in2.fq]\n"); return 1; }
	if ((rc = ssq_index_load(argv[optind], device, &idx))) die("ssq_index_load", rc);
	if (!(f1 = fq_open(argv[optind + 1]))) { fprintf(stderr, "[E::main_mem] fail to open file `%s'.\n", argv[optind + 1]); return 1; }
	if (optind + 2 < argc) {
		if (smart_pe) fprintf(stderr, "[W::main_mem] when '-p' is in use, the second query file is ignored.\n");
		else { if (!(f2 = fq_open(argv[optind + 2]))) { fprintf(stderr, "[E::main_mem] fail to open file `%s'.\n", argv[optind + 2]); return 1; } paired = 1; }
	}
	{ /* header: @SQ from the index, @RG as given, @PG with the command line */
		const int ns = (int)ssq_index_info(idx, 3);
		for (i = 0; i < ns; ++i) { int64_t len; const char *nm = ssq_index_contig(idx, i, &len); printf("@SQ\tSN:%s\tLN:%lld\n", nm, (long long)len); }
		if (rg_line) printf("%s\n", rg_line);
		printf("@PG\tID:bwa\tPN:bwa\tVN:%s\tCL:%s", SHIM_VERSION, prog);
		for (i = 0; i < argc; ++i) printf(" %s", argv[i]);
		printf("\n");
	}
	for (;;) {
		const long size = read_batch(chunk_size * n_threads, f1, f2, &r1, &r2, &v, keep_comment);
		char **out;
		read_t **se, **pe;
		int n_se = 0, n_pe = 0;
		if (v.n == 0) break;
		fprintf(stderr, "[M::process] read %d sequences (%ld bp)...\n", v.n, size);
		out = (char**)calloc(v.n, sizeof(char*));
		se = (read_t**)malloc(sizeof(read_t*) * v.n); pe = (read_t**)malloc(sizeof(read_t*) * v.n);
		if (smart_pe) { /* interleaved input: adjacent reads with equal names are mates, the others single-end */
			int has_last = 1;
			for (i = 1; i < v.n; ++i) {
				if (has_last) {
					if (strcmp(v.a[i].name, v.a[i - 1].name) == 0) { pe[n_pe++] = &v.a[i - 1]; pe[n_pe++] = &v.a[i]; has_last = 0; }
					else se[n_se++] = &v.a[i - 1];
				} else has_last = 1;
			}
			if (has_last) se[n_se++] = &v.a[v.n - 1];
			fprintf(stderr, "[M::process] %d single-end sequences; %d paired-end sequences\n", n_se, n_pe);
		} else if (paired) { for (i = 0; i < v.n; ++i) pe[n_pe++] = &v.a[i]; }
		else { for (i = 0; i < v.n; ++i) se[n_se++] = &v.a[i]; }
		if (n_se) run_sub(idx, &opt, se, n_se, n_processed, 0, 0, rg_id, out);
		if (n_pe) run_sub(idx, &opt, pe, n_pe, n_processed + n_se, 1, pes0, rg_id, out);
		n_processed += v.n;
		for (i = 0; i < v.n; ++i) {
			if (out[i]) { fputs(out[i], stdout); free(out[i]); }
			free(v.a[i].name); free(v.a[i].seq); free(v.a[i].qual); free(v.a[i].comment);
		}
		free(out); free(se); free(pe);
	}
	fflush(stdout);
	ssq_index_free(idx);
	return 0;
}

static int main_index(int argc, char **argv)
{
	const char *prefix = 0, *fa = 0;
	int i, rc, device = getenv("SSQ_DEVICE") ? atoi(getenv("SSQ_DEVICE")) : 0;
	for (i = 1; i < argc; ++i) {
		if (!strcmp(argv[i], "-p") && i + 1 < argc) prefix = argv[++i];
		else if (!strcmp(argv[i], "-a") && i + 1 < argc) ++i; /* construction algorithm: the result is the same */
		else if (argv[i][0] != '-' && !fa) fa = argv[i];
	}
	if (!fa) { fprintf(stderr, "Usage: bwa index [-p prefix] <in.fasta>\n"); return 1; }
	if ((rc = ssq_index_build(fa, prefix ? prefix : fa, device))) die("ssq_index_build", rc);
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "Usage: bwa <index|mem> [options]   (B200-native shim over libssq, behaviour of bwa %s)\n", SHIM_VERSION); return 1; }
	if (!strcmp(argv[1], "index")) return main_index(argc - 1, argv + 1);
	if (!strcmp(argv[1], "mem")) return main_mem(argc - 1, argv + 1, argv[0]);
	fprintf(stderr, "[main] unrecognized command '%s'\n", argv[1]);
	return 1;
}
