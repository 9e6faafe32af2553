#!/bin/bash
# Development tool: the default bench line (headline only) with the producer warps sleeping between polls of an empty
# slot (FDG_JIT_BULK_PSLEEP, ns) -- under bench.py's sustained load the device runs into its power cap, and instructions
# that do no work cost clock.  A/B on one box, default first and last.
for v in 0 200 1000 0; do
  echo "== FDG_JIT_BULK_PSLEEP=$v"
  FDG_JIT_BULK_PSLEEP=$v python bench.py --no-configs --no-e2e --no-cpu --no-torch-emitter --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['samples_per_s']/1e6,2),'M samples/s', d['clocks'])"
done
