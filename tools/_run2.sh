B="python bench.py --no-others --no-cpu-baseline --no-parity --steps 1 --warmup 3 --mode 0 --audio-channels 2 --variant exact --batch 4096 --blocks 4"
ncu --set full --clock-control none --import-source on -k regex:"k_rf_demod_iq|k_bpf_dual_packed" -s 6 -c 2 -o gpurun_out/r2s_packed_fir $B > gpurun_out/r2s_ncu.log 2>&1
tail -3 gpurun_out/r2s_ncu.log
