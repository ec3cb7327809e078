#!/bin/bash
# same-box A/B of the serpentine tile order (TFB_CH_REVERSE): TFB_SERPENTINE=0 switches it off
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-parity --no-extra"
for rep in 1 2; do for w in ${@:-kw_ga kw dw dw_full}; do for s in 1 0; do
  echo -n "serpentine=$s $w: "; TFB_SERPENTINE=$s timeout 120 $B --workload $w 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['sanity']['mass_balance_rel_err'], d['sanity']['bad_values'])"
done; done; done
