#!/bin/bash
# Round-end validation on one B200: GPU test suite, smoke, the default bench line, K1 on the config-3 pyramid (timing + one ncu capture),
# and ncu counters of the K8 / K9 filter kernels.  Every step has its own timeout; ncu reports are summarised on the box and deleted.
O=gpurun_out/${1:-final}
mkdir -p $O
timeout 480 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 180 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 480 python bench.py > $O/bench1.json 2> $O/bench1.err; echo "bench rc=$?"; tail -c 600 $O/bench1.err
timeout 200 python tools/k1_big_pyramid.py --config 3 --log2n 19 21 > $O/k1_big.txt 2>&1; echo "k1_big rc=$?"; cat $O/k1_big.txt | tail -3
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k1_ncc --launch-skip 3 --launch-count 1 -o $O/k1_c3 \
    python tools/k1_big_pyramid.py --config 3 --log2n 19 --steps 1 > $O/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
[ -f $O/k1_c3.ncu-rep ] && python profiles/summarize.py $O/k1_c3.ncu-rep > $O/k1_c3_summary.txt 2>&1; rm -f $O/k1_c3.ncu-rep
timeout 240 ncu --set full --clock-control none -k 'regex:k8_neighbor|k9_' --launch-count 12 -o $O/k8k9 \
    python tools/profile_kernels.py --iters 1 > $O/ncu_k8.log 2>&1; echo "ncu k8 rc=$?"
[ -f $O/k8k9.ncu-rep ] && python profiles/summarize.py $O/k8k9.ncu-rep > $O/k8k9_summary.txt 2>&1; rm -f $O/k8k9.ncu-rep
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref arm rc=$?"; tail -c 400 $O/bench_ref.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
