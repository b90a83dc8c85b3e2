ncu --set full --clock-control none --import-source on -k regex:"k_channelize2" -s 3 -c 1 -o gpurun_out/r2z_chan python bench.py --only-channelizer > gpurun_out/r2z_ncu.log 2>&1
tail -2 gpurun_out/r2z_ncu.log
