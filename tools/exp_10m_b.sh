for a in "--sort 3 --knobs 8,14,10,18" "--sort 2 --knobs 9,14,10,18" "--sort 2 --knobs 10,14,10,18" "--sort 2 --knobs 8,14,10,19" "--sort 2 --knobs 8,14,10,17" "--sort 2 --knobs 8,6,10,18" "--sort 2 --knobs 8,14,4,18" "--sort 2 --knobs 8,14,10,2"; do
  python tools/run_configs.py soup10m --rays 12500000 --check-rays 200000 $a 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$a', {k: d.get(k) for k in ('trace_ms_max_over_ranks', 'mrays_s', 'roofline_frac_per_gpu', 'sample_bit_identical_to_oracle')})"
done
