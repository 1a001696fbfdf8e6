#!/bin/bash
# Development: traversal variants on the 10 M-triangle heightfield (12.5 M random rays, one GPU).
for a in "--sort 0" "--sort 2" "--sort 3" "--sort 0 --knobs 12,14,10,18" "--sort 2 --knobs 12,14,10,18" "--sort 2 --knobs 8,14,10,42" "--sort 2 --knobs 8,14,10,18" "$@"; do
  python tools/run_configs.py soup10m --rays 12500000 --check-rays 200000 $a 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$a', {k: d.get(k) for k in ('trace_ms_max_over_ranks', 'mrays_s', 'sort_rays', 'roofline_frac_per_gpu', 'sample_bit_identical_to_oracle', 'gpu_build_ms', 'bytes_per_ray')})"
done
