set -x
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -5
B="python bench.py --no-others --no-cpu-baseline --steps 5 --warmup 3 --mode 0 --audio-channels 2"
$B --variant exact_scalar --batch 16384 --blocks 4 > gpurun_out/r2r_st_scalar_16k.json 2> gpurun_out/r2r.err
$B --variant exact --batch 16384 --blocks 4 > gpurun_out/r2r_st_packed_16k.json 2>> gpurun_out/r2r.err
$B --variant exact --batch 32768 --blocks 2 > gpurun_out/r2r_st_packed_32k.json 2>> gpurun_out/r2r.err
$B --variant exact --batch 65535 --blocks 1 > gpurun_out/r2r_st_packed_64k.json 2>> gpurun_out/r2r.err
python bench.py --no-others --no-cpu-baseline --steps 5 --warmup 3 --mode 2 --variant exact > gpurun_out/r2r_mono2_exact_packed.json 2>> gpurun_out/r2r.err
python bench.py --no-others --no-cpu-baseline --steps 5 --warmup 3 --mode 2 --variant exact_scalar > gpurun_out/r2r_mono2_exact_scalar.json 2>> gpurun_out/r2r.err
tail -3 gpurun_out/r2r.err
