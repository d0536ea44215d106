#!/bin/bash
# A/B of the host-buffer NCC call (bench.py `e2e`): the streamed K1 launch with different chunk sizes / copy streams against the older
# per-chunk launches, and the raw pinned H2D rate of the box for the same bytes.
# usage (on a GPU box): bash tools/e2e_sweep.sh > gpurun_out/e2e_sweep.txt
B="timeout 200 python bench.py --no-cpu-baseline --pipeline-iters 0 --steps 100"
run() {
  name="$1"; shift
  env "$@" $B 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('%-28s value %7.1f  e2e %7.1f M evals/s  %.3f ms/step' % ('$name', d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['ms_per_step']))
except Exception as e:
    print('%-28s failed: %s' % ('$name', e))"
}
python - <<'PY'
import torch
for nb, parts in ((62914560, 1), (62914560, 4), (62914560, 40), (8388608, 1)):
    h = torch.empty(nb, dtype=torch.uint8).pin_memory(); d = torch.empty(nb, dtype=torch.uint8, device="cuda")
    step = nb // parts
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for i in range(parts):
            d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print("raw pinned H2D %9d bytes in %2d copies: %.3f ms = %.1f GB/s" % (nb, parts, best, nb / best / 1e6))
PY
run "streamed 2^17 x2 (default)" PMK_X=0
run "streamed 2^17 x1" PMK_E2E_STREAMS=1
run "streamed 2^16 x2" PMK_E2E_CHUNK_LOG2=16
run "streamed 2^18 x2" PMK_E2E_CHUNK_LOG2=18
run "chunks (per-chunk launches)" PMK_E2E_MODE=chunks
