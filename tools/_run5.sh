python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo BENCH_EXIT $?
tail -2 gpurun_out/r2x_bench.err
