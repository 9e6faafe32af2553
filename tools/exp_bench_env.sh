#!/bin/bash
# sustained (power-capped) headline throughput of bench.py under sets of environment overrides, one per argument:
#   bash tools/exp_bench_env.sh "" "FDG_CSE_SCOPE=2000" "FDG_CSE_SCOPE=2000 FDG_JIT_RELOAD_GAP=2000"
for v in "$@"; do
  echo "== ${v:-default}"
  env $v python bench.py --no-e2e --no-cpu --no-torch-emitter --no-configs --steps 6 --warmup 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        j = d['config']['jit']
        print(round(d['samples_per_s'] / 1e6, 2), 'M samples/s', d['clocks']['sm_mhz'], 'MHz', j['kernels'], 'kernels', j['fp64_instr'], 'fp64', j.get('refetch_loads'), 'refetched')
"
done
