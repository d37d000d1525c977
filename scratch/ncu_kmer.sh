#!/bin/bash
# ncu --set full of the k-mer counting kernel on a 2 M-read shotgun batch (third launch: the table is warm) -> gpurun_out/$1.ncu-rep
ncu --set full --clock-control none --import-source on -k 'regex:k_kmer$' -s 2 -c 1 -o gpurun_out/$1 -f python scratch/kmer_bench.py > gpurun_out/$1.log 2>&1
