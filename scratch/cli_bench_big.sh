#!/bin/bash
# Steady-state wall-clock of the drop-in CLI on a larger uncompressed C2 input (tmpfs); reference timed on a 1/8 sample.
set -e
N=${1:-8000000}
D=/dev/shm/fqb; rm -rf $D; mkdir -p $D
python - <<PY
from faqcs_b200 import synth
w = synth.c2(1000000)
r1, r2 = bytes(w.r1), bytes(w.r2)
reps = $N // 1000000
with open("$D/r1.fastq","wb") as f:
    for _ in range(reps): f.write(r1)
with open("$D/r2.fastq","wb") as f:
    for _ in range(reps): f.write(r2)
open("$D/s1.fastq","wb").write(r1); open("$D/s2.fastq","wb").write(r2)
PY
ls -la $D | tail -4; free -g | head -2
export FAQCS_B200_TIMING=1
for mode in pieces bytes; do [ $mode = bytes ] && export FAQCS_B200_CLI_BYTES=1; echo "== output mode: $mode"
for mb in 64 256; do
t0=$(date +%s.%N); faqcs_b200/host/faqcs_b200 -1 $D/r1.fastq -2 $D/r2.fastq -d $D/gpu --prefix QC --trim_only --batch_mb $mb 2>&1 | grep timing; t1=$(date +%s.%N)
python -c "print(\"faqcs_b200 CLI, $N pairs, batch_mb $mb: %.2f s -> %.2f M reads/s\" % ($t1 - $t0, 2*$N/($t1 - $t0)/1e6))"
rm -rf $D/gpu
done
done
unset FAQCS_B200_CLI_BYTES
t0=$(date +%s.%N); oracle/_ref/FaQCs -1 $D/s1.fastq -2 $D/s2.fastq -d $D/ref --prefix QC --trim_only -t $(nproc) > $D/ref.log 2>&1 || true; t1=$(date +%s.%N)
python -c "print(\"reference FaQCs -t $(nproc), 1000000 pairs: %.2f s -> %.3f M reads/s\" % ($t1 - $t0, 2e6/($t1 - $t0)/1e6))"
