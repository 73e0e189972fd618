#!/bin/bash
# quick engine-option sweep on the GPU box: one JSON line per variant in gpurun_out/sweep.jsonl
# each argument: "[LIB=<variant name>] <bench.py flags>"
out=gpurun_out/sweep.jsonl; : > $out
for v in "$@"; do
  echo "## $v" >> $out
  lib=""; args="$v"
  if [[ "$v" == LIB=* ]]; then lib="${v%% *}"; lib="${lib#LIB=}"; args="${v#* }"; fi
  if [ -n "$lib" ]; then export AKR_B200_CUDA_LIB=$PWD/build/variants/libakari_b200_$lib.so; else unset AKR_B200_CUDA_LIB; fi
  python bench.py --steps 1 --warmup 1 --spp 128 --cpu-seconds 0.5 $args >> $out 2>> gpurun_out/sweep.err
done
