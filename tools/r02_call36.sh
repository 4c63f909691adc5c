F="python tools/layer_times.py 7 128"
$F > /dev/null 2>&1 || exit 1
for k in upsample2_sep_kernel head_tile_kernel input_prep_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -o /tmp/pw_$k $F > /dev/null 2>&1
  ncu -i /tmp/pw_$k.ncu-rep --page details > gpurun_out/r02_pw_$k.details.txt 2>/dev/null
  ncu -i /tmp/pw_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_pw_$k.sass.csv 2>/dev/null
done
# apply_sse instances: dc6 is the 2nd <16,1> launch of a forward, ec1 the <8,1>
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:apply_sse_kernel<\(int\)16' -s 5 -c 1 -o /tmp/pw_sse16 $F > /dev/null 2>&1
ncu -i /tmp/pw_sse16.ncu-rep --page details > gpurun_out/r02_pw_sse16.details.txt 2>/dev/null
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:apply_sse_kernel<\(int\)8' -s 2 -c 1 -o /tmp/pw_sse8 $F > /dev/null 2>&1
ncu -i /tmp/pw_sse8.ncu-rep --page details > gpurun_out/r02_pw_sse8.details.txt 2>/dev/null
ls -la gpurun_out/r02_pw_*
