// ssq_mem.cu — `bwa mem` for one batch through string arrays (upstream mem_process_seqs(); `$BWA mem`,
// /root/reference/bin/speedseq:438,468): a convenience form of the HBM-resident pipeline in ssq_pipe.cu for callers that hold
// their reads as separate C strings.  It gathers the strings into the concatenated layout ssq_aligner_run() takes, runs one
// batch with the samblaster stage switched off, and returns the records as one malloc'd string.  The CLI shim drives a
// persistent aligner object directly (no per-batch set-up); this entry point creates and frees one per call.
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "ssq_host.h"

extern "C" int ssq_mem_batch_sam(const ssq_index_t *idx, const ssq_opts_t *opt, int n_reads, const char *const *names, const char *const *seqs, const char *const *quals,
                                 const char *const *comments, int64_t n_processed, int paired, const ssq_pestat_t *pes0, const char *rg_id, int verbose, char **sam_out, size_t *sam_len, size_t *read_sam_off)
{
	if (!idx || !opt || n_reads < 0 || !sam_out || (n_reads && (!names || !seqs))) return SSQ_EINVAL;
	if (paired && (n_reads & 1)) { ssq_set_error("paired batch with an odd number of reads"); return SSQ_EINVAL; }
	std::string seqb, qualb, nameb, cmtb;
	std::vector<uint64_t> seq_off(n_reads + 1, 0);
	std::vector<uint32_t> name_off(n_reads + 1, 0), cmt_off(n_reads + 1, 0);
	bool has_qual = quals != 0;
	for (int i = 0; i < n_reads && has_qual; ++i) if (!quals[i]) has_qual = false;
	for (int i = 0; i < n_reads; ++i) {
		const size_t l = strlen(seqs[i]);
		seqb.append(seqs[i], l); seq_off[i + 1] = seqb.size();
		if (has_qual) { if (strlen(quals[i]) != l) { ssq_set_error("read %d: sequence and quality differ in length", i); return SSQ_EINVAL; } qualb.append(quals[i], l); }
		nameb += names[i]; name_off[i + 1] = (uint32_t)nameb.size();
		if (comments && comments[i]) cmtb += comments[i];
		cmt_off[i + 1] = (uint32_t)cmtb.size();
	}
	ssq_reads_t rd;
	memset(&rd, 0, sizeof rd);
	rd.n_reads = n_reads; rd.paired = paired; rd.seq = seqb.data(); rd.seq_off = seq_off.data(); rd.qual = has_qual ? qualb.data() : 0;
	rd.name = nameb.data(); rd.name_off = name_off.data(); rd.comment = comments ? cmtb.data() : 0; rd.comment_off = comments ? cmt_off.data() : 0; rd.n_processed = n_processed;
	ssq_aligner_t *al = 0;
	int rc = ssq_aligner_create(idx, opt, 0, rg_id, &al);
	if (rc) return rc;
	ssq_sam_t out;
	rc = ssq_aligner_run(al, &rd, pes0, verbose, &out);
	if (!rc) {
		*sam_out = (char*)malloc(out.len[0] + 1);
		if (!*sam_out) rc = SSQ_ENOMEM;
		else {
			memcpy(*sam_out, out.text[0], out.len[0]); (*sam_out)[out.len[0]] = 0;
			if (sam_len) *sam_len = out.len[0];
			if (read_sam_off) for (int i = 0; i <= n_reads; ++i) read_sam_off[i] = (size_t)out.read_off[i];
		}
	}
	ssq_aligner_free(al);
	return rc;
}

extern "C" void ssq_free(void *p) { free(p); }
