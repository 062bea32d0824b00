#!/bin/bash
# compute-sanitizer over the whole product path on a small input: the `bwa` shim (fused samblaster stage, so every pipeline kernel
# runs) on the first 600 example reads against the example reference.  memcheck, then racecheck (shared-memory hazards of the
# warp kernels).  usage: tools/sanitize_cli.sh [workdir]   (needs a GPU)
ROOT=$(cd "$(dirname "$0")/.." && pwd); W=${1:-/tmp/ssq_sanitize}; mkdir -p "$W"
gzip -dc "$ROOT/tests/golden/ex_ref.fa.gz" > "$W/ref.fa" 2>/dev/null || cp "$ROOT/oracle/_ref/stage/example/data/human_g1k_v37_20_42220611-42542245.fasta" "$W/ref.fa"
BWA=$ROOT/speedseq_b200/bin/bwa
timeout 120 "$BWA" index "$W/ref.fa" > "$W/index.log" 2>&1 || { echo "index failed"; tail -3 "$W/index.log"; exit 1; }
N=${SSQ_SANITIZE_READS:-600}; gzip -dc "$ROOT/tests/golden/ex_reads_2k.fq.gz" | head -n $((4 * N)) > "$W/reads.fq"
export SSQ_FUSE_SAMBLASTER="--excludeDups --addMateTags --maxSplitCount 2 --minNonOverlap 20"
timeout 60 "$BWA" mem -p "$W/ref.fa" "$W/reads.fq" 2>/dev/null | md5sum > "$W/plain.md5"
rc=0
for tool in ${SSQ_SANITIZE_TOOLS:-memcheck racecheck}; do
	timeout ${SSQ_SANITIZE_TIMEOUT:-100} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 --log-file "$W/$tool.log" "$BWA" mem -p "$W/ref.fa" "$W/reads.fq" > "$W/$tool.out" 2> "$W/$tool.err"
	r=$?
	md5sum < "$W/$tool.out" > "$W/$tool.md5"
	echo "== $tool: exit $r, output $(cmp -s "$W/plain.md5" "$W/$tool.md5" && echo identical to the uninstrumented run || echo DIFFERS)"
	grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|Warning:" "$W/$tool.log" | sort | uniq -c | sort -rn | head -8
	[ $r -ne 0 ] && { rc=1; head -c 1500 "$W/$tool.err"; head -c 3000 "$W/$tool.log"; }
done
exit $rc
