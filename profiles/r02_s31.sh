#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
for so in athenak_b200/libathenak_b200.so athenak_b200/var/lib_us.so athenak_b200/var/lib_us1.so; do
  echo "== $so"
  AKB_LIB_LAX=1 AKB_TRACE=1 ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -2
  AKB_LIB_LAX=1 AKB_TRACE=1 ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 35 2>&1 | grep -v "^$" | tail -2
done
