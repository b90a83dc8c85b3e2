#!/bin/bash
# usage: tools/bench_kernels.sh <mode> <variant> [extra bench args]  -> one line: value + per-kernel ms
m=$1; v=$2; shift 2
timeout 300 python bench.py --mode $m --variant $v --no-cpu-baseline --no-others --no-parity --e2e-blocks 1 "$@" 2>&1 | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
print(d["config"]["mode"], d["config"]["variant"][:5], "GS/s", round(d["value"] / 1e3, 1), "step ms", round(d["ms_per_step"], 4),
      {k: round(v, 4) for k, v in d["roofline"]["kernel_ms_per_step"].items()})'
