#!/bin/bash
# quick engine-option sweep on the GPU box: writes one JSON line per variant to gpurun_out/sweep.jsonl
out=gpurun_out/sweep.jsonl; : > $out
run() { echo "## $*" >> $out; python bench.py --steps 1 --warmup 1 --spp 128 --cpu-seconds 0.5 "$@" >> $out 2>> gpurun_out/sweep.err; }
for tm in 1 2; do for wave in 1048576 4194304 16777216 67108864; do run --trace-mode $tm --wave $wave; done; done
run --trace-mode 2 --sort 2 --wave 4194304
