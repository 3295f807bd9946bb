#!/bin/bash
mkdir -p gpurun_out
python profiles/pcie_duplex.py 2>&1 | tail -8
for so in athenak_b200/libathenak_b200.so athenak_b200/var/lib_g0.so; do
  echo "== $so"
  AKB_TRACE=1 ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -2
done
echo "== main, no sweep streams"; AKB_SWEEP_STREAMS=0 python profiles/bench_variants.py --only headline --steps 10 2>&1 | tail -1
echo "== main, no overlap U"; AKB_OVERLAP_U=0 python profiles/bench_variants.py --only headline --steps 10 2>&1 | tail -1
