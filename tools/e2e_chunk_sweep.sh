#!/bin/bash
# e2e (host-pointer pipeline) step time against the chunk size of lteb200_dl_receive_host; writes gpurun_out/e2e_chunk_sweep.txt
mkdir -p gpurun_out
: > gpurun_out/e2e_chunk_sweep.txt
for c in 125 250 500 1000 2000 4000; do
  LTEB200_E2E_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('chunk', $c, 'device ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e sf/s', d['e2e']['value'])" >> gpurun_out/e2e_chunk_sweep.txt
done
cat gpurun_out/e2e_chunk_sweep.txt
