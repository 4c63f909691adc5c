# ncu source-level stall profile of sse_bwd_a (dc5: C = 32) with and without the bulk-copy ring
python -c "import __graft_entry__ as g; g.build()"
T="python tools/time_train.py 8 128"
cap() { # name env
  env $2 ncu --set full --clock-control none --import-source on -k regex:sse_bwd_a_kernel -s 1 -c 1 -o /tmp/c63_$1 $T > /dev/null 2>&1
  ncu -i /tmp/c63_$1.ncu-rep --page details > gpurun_out/r02_c63_$1.details.txt 2>/dev/null
  ncu -i /tmp/c63_$1.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_c63_$1.sass.csv 2>/dev/null
}
cap ring SEUNET_BWDA_RING=1
cap noring SEUNET_BWDA_RING=0
ls -la gpurun_out/r02_c63*
