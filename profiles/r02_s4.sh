#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tail.py -m gpu -x -q > gpurun_out/r02_s4_tail.log 2>&1
tail -25 gpurun_out/r02_s4_tail.log
for so in athenak_b200/var/lib_g0.so athenak_b200/libathenak_b200.so; do
  echo "== $so"
  ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 10
  AKB_TAIL=0 ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 10
done
for tag in y g0; do
  so=athenak_b200/libathenak_b200.so; [ $tag = g0 ] && so=athenak_b200/var/lib_g0.so
  AKB_TAIL=0 ATHENAK_B200_LIB=$PWD/$so ncu --set full --import-source on --clock-control none -k regex:flux_ -s 12 -c 3 -o gpurun_out/r02_s4_$tag -f \
     python profiles/bench_variants.py --only headline --steps 2 > gpurun_out/r02_s4_ncu_$tag.log 2>&1
done
