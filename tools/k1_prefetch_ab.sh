#!/bin/bash
# A/B of K1's L2 prefetch (PMK_K1_PREFETCH=0/1) on a pyramid larger than L2 (config 3) and on the L2-resident config 2,
# the K1 parity tests with the prefetch forced on, then the whole GPU suite with the default (automatic) setting.
O=gpurun_out/${1:-ab}
mkdir -p $O
for pf in 0 1; do
  PMK_K1_PREFETCH=$pf timeout 150 python tools/k1_big_pyramid.py --config 3 --log2n 19 21 > $O/k1_c3_pf$pf.txt 2>&1
  echo "config 3 prefetch $pf rc=$?"; tail -2 $O/k1_c3_pf$pf.txt
done
PMK_K1_PREFETCH=1 timeout 120 python -m pytest tests/test_k1_gpu.py tests/test_golden.py -m gpu -q -x > $O/pytest_k1_pf1.log 2>&1; echo "pytest pf1 rc=$?"; tail -2 $O/pytest_k1_pf1.log
for pf in 0 1; do
  PMK_K1_PREFETCH=$pf timeout 60 python tools/k1_big_pyramid.py --config 2 --log2n 20 > $O/k1_c2_pf$pf.txt 2>&1; echo "config 2 prefetch $pf rc=$?"; tail -1 $O/k1_c2_pf$pf.txt
done
timeout 300 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
