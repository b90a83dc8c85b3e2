#!/bin/bash
# usage: tools/bench_stereo.sh <batch> [mode]  -> stereo throughput and per-kernel ms (4 blocks per capture)
b=$1; m=${2:-0}
timeout 300 python bench.py --mode $m --audio-channels 2 --batch $b --blocks 4 --no-cpu-baseline --no-others --no-parity --e2e-blocks 1 2>&1 | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
print(d["config"]["workload"][:48], round(d["value"] / 1e3, 1), "GS/s", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})'
