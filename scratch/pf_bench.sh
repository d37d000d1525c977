#!/bin/bash
bash scratch/variant_bench.sh base nopf fpf4 fpf1
for t in 148 592 1184; do
  echo "emit prefetch tiles $t"; FAQCS_B200_EMIT_PREFETCH_TILES=$t bash scratch/variant_bench.sh base
done
