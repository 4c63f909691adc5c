for v in 1048576 2097152 4194304 33554432; do for b in 1 8; do echo "== conc_vox $v B=$b"; SEUNET_BWD_CONC_VOX=$v timeout 300 python tools/time_train.py $b 128 2>&1 | head -1; done; done
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py tests/test_gpu_wgrad.py -x -q 2>&1 | tail -3
