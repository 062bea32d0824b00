/* ssq_fuse.h — the private framing between the `bwa` and `samblaster` shims when samblaster's stage runs fused inside `bwa mem`
 * on the device (SSQ_FUSE_SAMBLASTER, see bwa_main.c / INTEGRATION.md).  After the SAM header `bwa` writes one marker line
 *     @CO\tssq-fused-v1\t<option description>
 * and then frames: 8-byte magic, u64 stream (0 main records, 1 splitters, 2 discordants), u64 payload length, payload (complete
 * SAM records).  `samblaster` drops the marker line, refuses to continue when its own argv describes different options (the
 * records were selected and flagged under the options `bwa` was given), and copies each payload to stdout / --splitterFile /
 * --discordantFile.  Nothing of this reaches the pipeline's consumers (speedseq:440-448).
 *
 * BAM mode (SSQ_FUSE_BAM=1 on top of the above; the marker line ends in "\tbam"): the main records do not travel as text at all.
 * `bwa` sends each batch's main stream as stream 3 = one run of BAM records, coordinate-sorted on the device; `samblaster` writes
 * the header text, the line SSQ_BAM_RUNS_MARKER and then those frames unchanged to stdout; the `sambamba` shim (sambamba_main.c)
 * recognises that on the stdin of `view -S -f bam` (passes it on) and of `sort` (merges the runs, rewrites the header the way
 * sambamba does, writes the BGZF file named by -o).  The two side streams stay SAM text: speedseq:443,446 run them through gawk. */
#ifndef SSQ_FUSE_H
#define SSQ_FUSE_H
#include <stdint.h>
#include <stdio.h>
#define SSQ_FUSE_MARKER "@CO\tssq-fused-v1\t"
#define SSQ_FRAME_MAGIC "SSQFRAME"
#define SSQ_BAM_RUNS_MARKER "@CO\tssq-bam-runs-v1\n"
#define SSQ_STREAM_BAM_RUN 3
typedef struct { char magic[8]; uint64_t stream, len; } ssq_frame_hdr_t;
static inline void ssq_fuse_describe(char *buf, size_t cap, int exclude_dups, int add_mate_tags, int remove_dups, int max_split_count, int min_non_overlap, int min_indel_size, int max_unmapped_bases)
{
	snprintf(buf, cap, "excludeDups=%d addMateTags=%d removeDups=%d maxSplitCount=%d minNonOverlap=%d minIndelSize=%d maxUnmappedBases=%d", !!exclude_dups, !!add_mate_tags, !!remove_dups,
	         max_split_count, min_non_overlap, min_indel_size, max_unmapped_bases);
}
#endif
