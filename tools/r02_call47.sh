timeout 300 python tools/window_step.py 2 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'apply_|upsample2|head_tile|input_prep|hu_windows|finalize|headw|clear' -o /tmp/c47_hbm python tools/window_step.py 2 > gpurun_out/r02_c47_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_hbm_table.py /tmp/c47_hbm.ncu-rep > gpurun_out/r02_c47_hbm_table.txt 2>&1; tail -45 gpurun_out/r02_c47_hbm_table.txt
cp /tmp/c47_hbm.ncu-rep gpurun_out/r02_c47_hbm.ncu-rep 2>/dev/null; ls -la gpurun_out/r02_c47_hbm.ncu-rep
