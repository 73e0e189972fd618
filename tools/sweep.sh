#!/bin/bash
# quick engine-option sweep on the GPU box: writes one JSON line per variant to gpurun_out/sweep.jsonl
out=gpurun_out/sweep.jsonl; : > $out
run() { echo "## $*" >> $out; python bench.py --steps 1 --warmup 1 --spp 128 --cpu-seconds 0.5 "$@" >> $out 2>> gpurun_out/sweep.err; }
for v in "$@"; do run $v; done
