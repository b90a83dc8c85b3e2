python -m pytest tests/test_gpu_round2.py tests/test_gpu_fast.py -x -q -m gpu 2>&1 | tail -3
python bench.py --no-others --no-cpu-baseline --steps 10 --warmup 3 --mode 3 > gpurun_out/r2y_mode3_tc.json 2> gpurun_out/r2y.err
tail -2 gpurun_out/r2y.err
