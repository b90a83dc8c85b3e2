# usage: bash tools/exp_rf_variants.sh v1 v2 ...   (SDR_RF_VARIANT values)
for v in "$@"; do
  echo "=== variant $v"
  SDR_RF_VARIANT=$v python -m pytest tests/test_gpu_pipeline.py -q -x -k "matches_oracle and (F-1-0 or F-1-2 or F-2-0)" 2>&1 | tail -1
  SDR_RF_VARIANT=$v python bench.py --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f MS/s'%d['value'], d['roofline']['kernel_ms_per_step'])"
done
