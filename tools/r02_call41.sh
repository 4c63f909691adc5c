timeout 600 python tools/ab_window.py 3 > gpurun_out/r02_c41_ab.txt 2>&1; cat gpurun_out/r02_c41_ab.txt
