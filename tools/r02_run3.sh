python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -5
GRIDS="512" VARS="2 4 5 6 7" LZS="32" bash tools/pair_ab.sh 2>&1 | tee gpurun_out/pair_ab3.log
GRIDS="256" VARS="2 4 5" LZS="16 32" bash tools/pair_ab.sh 2>&1 | tee -a gpurun_out/pair_ab3.log
python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=4 pair_lz=32 > gpurun_out/plain_pair4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_pair -s 4 -c 2 -o gpurun_out/prof_pair_v4 -f python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=4 pair_lz=32 > gpurun_out/ncu_pair_v4.log 2>&1
# multi-GPU pair path cannot run here (1 GPU); 
