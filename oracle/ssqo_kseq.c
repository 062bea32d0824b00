/* ssqo_kseq.c — ORACLE (test infrastructure). See ssqo_kseq.h. */
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <zlib.h>
#include "ssqo_kseq.h"

#define BUFSZ 65536
typedef struct { gzFile fp; unsigned char *buf; int beg, end, eof; } stream_t;

static int st_getc(stream_t *s)
{
	if (s->eof && s->beg >= s->end) return -1;
	if (s->beg >= s->end) {
		s->beg = 0;
		s->end = gzread(s->fp, s->buf, BUFSZ);
		if (s->end <= 0) { s->eof = 1; s->end = 0; return -1; }
	}
	return s->buf[s->beg++];
}

static void str_push(ssqo_str_t *v, int c)
{
	if (v->l + 2 > v->m) { v->m = v->m ? v->m << 1 : 64; v->s = (char*)realloc(v->s, v->m); }
	v->s[v->l++] = (char)c; v->s[v->l] = 0;
}

/* delim: 0 = whitespace, 2 = newline. append (or not) until delimiter; returns delimiter or -1 */
static int st_getuntil(stream_t *s, int delim, ssqo_str_t *v, int append)
{
	int c, got = 0;
	if (!append) { v->l = 0; if (v->s) v->s[0] = 0; }
	while ((c = st_getc(s)) >= 0) {
		got = 1;
		if (delim == 2 ? c == '\n' : isspace(c)) break;
		str_push(v, c);
	}
	if (!got && c < 0) return -1;
	if (v->s == 0) str_push(v, 0), v->l = 0;
	if (delim == 2 && v->l > 0 && v->s[v->l - 1] == '\r') v->s[--v->l] = 0;
	return c < 0 ? 0 : c;
}

ssqo_kseq_t *ssqo_kseq_open(const char *fn)
{
	ssqo_kseq_t *ks;
	stream_t *s;
	gzFile fp = strcmp(fn, "-") == 0 ? gzdopen(0, "r") : gzopen(fn, "r");
	if (!fp) return 0;
	ks = (ssqo_kseq_t*)calloc(1, sizeof(*ks));
	s = (stream_t*)calloc(1, sizeof(*s));
	s->fp = fp; s->buf = (unsigned char*)malloc(BUFSZ);
	ks->stream = s;
	return ks;
}

void ssqo_kseq_close(ssqo_kseq_t *ks)
{
	stream_t *s;
	if (!ks) return;
	s = (stream_t*)ks->stream;
	gzclose(s->fp); free(s->buf); free(s);
	free(ks->name.s); free(ks->comment.s); free(ks->seq.s); free(ks->qual.s);
	free(ks);
}

int ssqo_kseq_read(ssqo_kseq_t *ks)
{
	stream_t *s = (stream_t*)ks->stream;
	int c;
	if (ks->last_char == 0) { /* jump to the next header line */
		while ((c = st_getc(s)) != -1 && c != '>' && c != '@');
		if (c == -1) return -1;
		ks->last_char = c;
	}
	ks->comment.l = ks->seq.l = ks->qual.l = 0;
	if (ks->comment.s) ks->comment.s[0] = 0;
	if ((c = st_getuntil(s, 0, &ks->name, 0)) < 0) return -1;
	if (c != '\n') st_getuntil(s, 2, &ks->comment, 0);
	if (ks->seq.s == 0) { ks->seq.m = 256; ks->seq.s = (char*)malloc(ks->seq.m); ks->seq.s[0] = 0; }
	while ((c = st_getc(s)) != -1 && c != '>' && c != '+' && c != '@') {
		if (c == '\n') continue;
		str_push(&ks->seq, c);
		st_getuntil(s, 2, &ks->seq, 1);
	}
	if (c == '>' || c == '@') ks->last_char = c;
	ks->seq.s[ks->seq.l] = 0;
	if (c != '+') { if (c == -1) ks->last_char = 0; return (int)ks->seq.l; }
	if (ks->qual.m < ks->seq.m) { ks->qual.m = ks->seq.m; ks->qual.s = (char*)realloc(ks->qual.s, ks->qual.m); }
	while ((c = st_getc(s)) != -1 && c != '\n'); /* skip the rest of the '+' line */
	if (c == -1) return -2;
	if (ks->qual.s) ks->qual.s[0] = 0;
	while (st_getuntil(s, 2, &ks->qual, 1) >= 0 && ks->qual.l < ks->seq.l);
	ks->last_char = 0;
	if (ks->seq.l != ks->qual.l) return -2;
	return (int)ks->seq.l;
}
