#!/bin/bash
# Wall-clock of the drop-in CLI vs the reference binary on the same uncompressed C2 files (tmpfs).
set -e
N=${1:-1000000}
D=/dev/shm/fqb; rm -rf $D; mkdir -p $D
python - <<PY
from faqcs_b200 import synth
w = synth.c2($N)
open("$D/r1.fastq","wb").write(bytes(w.r1)); open("$D/r2.fastq","wb").write(bytes(w.r2))
PY
ls -la $D
nproc
export OMP_NUM_THREADS=$(nproc)
/usr/bin/time -v true 2>/dev/null || true
export FAQCS_B200_TIMING=1
t0=$(date +%s.%N); faqcs_b200/host/faqcs_b200 -1 $D/r1.fastq -2 $D/r2.fastq -d $D/gpu --prefix QC --trim_only 2>&1 | grep timing; t1=$(date +%s.%N)
python -c "print(\"faqcs_b200 CLI: %.2f s\" % ($t1 - $t0))"
t0=$(date +%s.%N); faqcs_b200/host/faqcs_b200 -1 $D/r1.fastq -2 $D/r2.fastq -d $D/gpu2 --prefix QC --trim_only 2>&1 | grep timing; t1=$(date +%s.%N)
python -c "print(\"faqcs_b200 CLI (2nd run): %.2f s\" % ($t1 - $t0))"
if [ -x oracle/_ref/FaQCs ]; then
t0=$(date +%s.%N); oracle/_ref/FaQCs -1 $D/r1.fastq -2 $D/r2.fastq -d $D/ref --prefix QC --trim_only -t $(nproc) > $D/ref.log 2>&1 || true; t1=$(date +%s.%N)
python -c "print(\"reference FaQCs -t $(nproc): %.2f s\" % ($t1 - $t0))"
cmp $D/gpu/QC.1.trimmed.fastq $D/ref/QC.1.trimmed.fastq && cmp $D/gpu/QC.2.trimmed.fastq $D/ref/QC.2.trimmed.fastq && cmp $D/gpu/QC.stats.txt $D/ref/QC.stats.txt && echo "outputs identical"
fi
