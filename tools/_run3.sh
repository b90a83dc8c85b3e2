for m in 0 1 3; do for f in 0 1 5 6; do
if [ $m != 0 ] && [ $f = 1 -o $f = 6 ]; then continue; fi
SDR_EXP_FORM=$f python bench.py --no-others --no-cpu-baseline --steps 5 --warmup 3 --mode $m --audio-channels 1 --variant exact --batch 4096 --blocks 4 > gpurun_out/r2w_m${m}_form$f.json 2>> gpurun_out/r2w.err
done; done
tail -3 gpurun_out/r2w.err
