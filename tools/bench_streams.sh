#!/bin/bash
# usage: tools/bench_streams.sh  -> stereo and mixed workloads over 1..8 pipeline handles / CUDA streams
run() { timeout 300 python bench.py "$@" --no-cpu-baseline --no-others --no-parity 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
print(d["config"]["workload"][:100], "|", round(d["value"] / 1e3, 1), "GS/s", round(d["ms_per_step"], 2), "ms", round(100 * d["roofline"]["whole_step_frac"], 2), "% HBM")'; }
for s in 1 2 4; do run --mode 0 --audio-channels 2 --batch 1024 --blocks 4 --steps 10 --streams $s; done
for s in 1 2 4 8; do run --mode 0 --audio-channels 2 --batch 16384 --blocks 4 --steps 5 --streams $s; done
for s in 2 4 8; do run --mode 0 --mixed --batch 8192 --blocks 4 --steps 5 --streams $s; done
for s in 1 2; do run --steps 20 --streams $s; done
