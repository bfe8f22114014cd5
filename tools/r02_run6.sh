./tools/tma_probe.bin 2>&1 | tee gpurun_out/tma_probe.log
python -m pytest tests/test_gpu_parity.py -x -q -k "convergence" 2>&1 | tail -15
python -m pytest tests/test_gpu_host_program.py -x -q 2>&1 | tail -5
cd /tmp && for i in 1 2; do /usr/bin/time -f "wall %e s" $GRAFT_REPO_ROOT/cfd-codes_b200/lbm3d_cavity --SAVETXT 0 | tail -4; done; cd $GRAFT_REPO_ROOT
