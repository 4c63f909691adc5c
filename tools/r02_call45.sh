timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -x -q 2>&1 | tail -3
for v in 0 32768 262144 1048576 2097152; do echo "B=1 conc_vox=$v"; SEUNET_BWD_CONC_VOX=$v timeout 300 python tools/time_train.py 1 128 2>&1 | head -1; done
for v in 0 262144 2097152 16777216; do echo "B=8 conc_vox=$v"; SEUNET_BWD_CONC_VOX=$v timeout 300 python tools/time_train.py 8 128 2>&1 | head -1; done
for v in 0 262144 1048576 4194304; do echo "B=2 conc_vox=$v"; SEUNET_BWD_CONC_VOX=$v timeout 300 python tools/time_train.py 2 128 2>&1 | head -1; done
