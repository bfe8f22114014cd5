for w in 8 4 2; do
python tools/slab_probe.py 512 $w 1 6 > gpurun_out/slab_plain_$w.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/slab_launches_$w.csv python tools/slab_probe.py 512 $w 1 6 > gpurun_out/slab_ncu_$w.log 2>&1
cat gpurun_out/slab_plain_$w.log
done
python -c "import __graft_entry__ as g; g.smoke()"
