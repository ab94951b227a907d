#!/bin/bash
cd "$(dirname "$0")/.."
for nt in 2 8 16; do
  KDSB200_WRITE_THREADS=$nt python tools/run_configs.py --configs 4 --scale 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads', $nt, 'e2e_s', d['e2e_s'], 'open', d['open_source_s'], 'pps', d['e2e_particles_per_s'])"
done
