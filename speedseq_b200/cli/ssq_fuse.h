/* ssq_fuse.h — the private framing between the `bwa` and `samblaster` shims when samblaster's stage runs fused inside `bwa mem`
 * on the device (SSQ_FUSE_SAMBLASTER, see bwa_main.c / INTEGRATION.md).  After the SAM header `bwa` writes one marker line
 *     @CO\tssq-fused-v1\t<option description>
 * and then frames: 8-byte magic, u64 stream (0 main records, 1 splitters, 2 discordants), u64 payload length, payload (complete
 * SAM records).  `samblaster` drops the marker line, refuses to continue when its own argv describes different options (the
 * records were selected and flagged under the options `bwa` was given), and copies each payload to stdout / --splitterFile /
 * --discordantFile.  Nothing of this reaches the pipeline's consumers (speedseq:440-448). */
#ifndef SSQ_FUSE_H
#define SSQ_FUSE_H
#include <stdint.h>
#include <stdio.h>
#define SSQ_FUSE_MARKER "@CO\tssq-fused-v1\t"
#define SSQ_FRAME_MAGIC "SSQFRAME"
typedef struct { char magic[8]; uint64_t stream, len; } ssq_frame_hdr_t;
static inline void ssq_fuse_describe(char *buf, size_t cap, int exclude_dups, int add_mate_tags, int remove_dups, int max_split_count, int min_non_overlap, int min_indel_size, int max_unmapped_bases)
{
	snprintf(buf, cap, "excludeDups=%d addMateTags=%d removeDups=%d maxSplitCount=%d minNonOverlap=%d minIndelSize=%d maxUnmappedBases=%d", !!exclude_dups, !!add_mate_tags, !!remove_dups,
	         max_split_count, min_non_overlap, min_indel_size, max_unmapped_bases);
}
#endif
