python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_a.log; cat gpurun_out/r02_pytest_gpu_a.log
bash tools/pair_ab.sh 2>&1 | tee gpurun_out/pair_ab2.log
python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=1 pair_lz=32 > gpurun_out/plain_pair.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_pair -s 4 -c 2 -o gpurun_out/prof_pair_v1 -f python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=1 pair_lz=32 > gpurun_out/ncu_pair_v1.log 2>&1
python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=0 pair_lz=32 > gpurun_out/plain_pair0.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_pair -s 4 -c 2 -o gpurun_out/prof_pair_v0 -f python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=0 pair_lz=32 > gpurun_out/ncu_pair_v0.log 2>&1
