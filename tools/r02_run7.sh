./tools/tma_probe.bin 2>&1 | tee gpurun_out/tma_probe.log
python -m pytest tests/test_gpu_parity.py -x -q -k "convergence" 2>&1 | tail -5
python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -5
GRIDS="512" VARS="7 8 9 10" LZS="32" bash tools/pair_ab.sh 2>&1 | tee gpurun_out/pair_ab7.log
GRIDS="512" VARS="8" LZS="16 64" bash tools/pair_ab.sh 2>&1 | grep -v '"pair": "0"' | tee -a gpurun_out/pair_ab7.log
GRIDS="256" VARS="7 8" LZS="16 32" bash tools/pair_ab.sh 2>&1 | tee -a gpurun_out/pair_ab7.log
cd /tmp; for i in 1 2; do S=$(date +%s.%N); $GRAFT_REPO_ROOT/cfd-codes_b200/lbm3d_cavity --SAVETXT 0 | tail -4; E=$(date +%s.%N); echo "default config wall: $(echo "$E - $S" | bc) s"; done; cd $GRAFT_REPO_ROOT
python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=8 pair_lz=32 > gpurun_out/plain_pair8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_pair -s 4 -c 2 -o gpurun_out/prof_pair_v8 -f python tools/time_lib.py cfd-codes_b200/liblbm_b200.so 512 pair=1 pair_variant=8 pair_lz=32 > gpurun_out/ncu_pair_v8.log 2>&1
