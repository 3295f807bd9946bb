#!/bin/bash
# marked-face repeat for the 3-point marches (PLM / donor cell) as well: full GPU suite, then the side table of the other families
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_s50_pytest.log
cat gpurun_out/r02_s50_pytest.log
python profiles/bench_variants.py --steps 10 > gpurun_out/r02_variants.txt 2> gpurun_out/r02_variants.err
python profiles/bench_variants.py --steps 6 --only "config 5" >> gpurun_out/r02_variants.txt 2>> gpurun_out/r02_variants.err
cat gpurun_out/r02_variants.txt; tail -2 gpurun_out/r02_variants.err
