for f in 1; do
SDR_EXP_FORM=$f ncu --set full --clock-control none --import-source on -k regex:"k_rf_demod_iq" -s 3 -c 1 -o gpurun_out/r2v_form$f python bench.py --no-others --no-cpu-baseline --no-parity --steps 1 --warmup 3 --mode 0 --audio-channels 1 --variant exact --batch 2048 --blocks 4 > gpurun_out/r2v_ncu$f.log 2>&1
done
tail -2 gpurun_out/r2v_ncu1.log
