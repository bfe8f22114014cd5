python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -5
GRIDS="512" VARS="7 8 9 10" LZS="32" bash tools/pair_ab.sh 2>&1 | tee gpurun_out/pair_ab5.log
GRIDS="512" VARS="8" LZS="16 64" bash tools/pair_ab.sh 2>&1 | grep -v '"pair": "0"' | tee -a gpurun_out/pair_ab5.log
GRIDS="256" VARS="7 8" LZS="16 32" bash tools/pair_ab.sh 2>&1 | tee -a gpurun_out/pair_ab5.log
for v in 7 8; do
python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=$v pair_lz=32 > gpurun_out/plain_pair$v.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_pair -s 4 -c 2 -o gpurun_out/prof_pair_v$v -f python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=$v pair_lz=32 > gpurun_out/ncu_pair_v$v.log 2>&1
done
