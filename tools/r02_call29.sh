for kb in 50 70 100 110; do echo "== stage cap $kb KB"; SEUNET_WG_STAGE_KB=$kb timeout 300 python tools/time_train.py 8 128 2>&1 | head -2; done
echo "== B=1"; DETAIL=1 timeout 300 python tools/time_train.py 1 128 > gpurun_out/r02_train_b1_kh.txt 2>&1; head -3 gpurun_out/r02_train_b1_kh.txt
