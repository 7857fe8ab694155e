#!/bin/bash
# A/B on the batch benchmark (bench.py --no-sharded): the L2 fetch-granularity hint (applied at pcgpu_init) and the number of
# threads of the pair-round kernel (PCGPU_BATCH_TDIV = 2: two blocks per SM, so that the pair kernels of two pipelines
# are co-resident on every SM and the DRAM-bound pass 1 / ALU-bound inversion of one overlap the multiply-bound pass 2 of the other)
cd "$(dirname "$0")/../.."
for g in 128 64 32; do
  for t in 1 2 3; do
    PCGPU_L2_FETCH_GRANULARITY=$g PCGPU_BATCH_TDIV=$t python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('l2_fetch', $g, 'tdiv', $t, 'value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'single_call_ms', round(d['single_call_ms_per_step'], 2), 'pair0_ms', round(d['roofline']['launch_ms'], 3), 'pair_rounds_ms', round(d['stage_ms_per_step']['affine_pair_rounds'], 2))"
  done
done
