#!/bin/bash
# round-2 call 1: parity of the gen-2 fast path, A/B timings, ncu of the shipped kernel
mkdir -p gpurun_out/r2c1
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
BAM_K=4096,512 timeout 300 python scratch/time_k1.py > $O/time_main.log 2>&1
for v in g1 ilp4 ilp6 e1c minb3; do
  BAM_K=4096,512 BAM_LIB=scratch/variants/libbam_gpu_$v.so timeout 300 python scratch/time_k1.py > $O/time_$v.log 2>&1
done
cat $O/time_*.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logpost_baratin_cb2 -c 1 -s 3 -o $O/cb2 python scratch/time_k1.py > $O/ncu.log 2>&1
tail -3 $O/pytest.log
