#!/bin/bash
# same-box A/B of the extended packed kernels: the in-tree build against a saved variant
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-parity --no-extra"
run() { echo -n "lib=${1##*/} $2: "; TFB_LIB_PATH=$1 timeout 120 $B --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['sanity']['mass_balance_rel_err'], d['sanity']['bad_values'])"; }
D=topoflow36_b200
for rep in 1 2; do
for w in kw_grid dw_grid; do
run $D/libtfb_b200.so $w; for v in ${@:-$D/_variants/libtfb_before_ext.so}; do run $v $w; done
done; done
