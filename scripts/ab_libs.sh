#!/bin/bash
# same-box A/B of kernel variants: usage: ab_libs.sh "<workloads>" lib1.so lib2.so ...   (2 repetitions)
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-parity --no-extra"
W="$1"; shift
run() { echo -n "lib=${1##*/} $2: "; TFB_LIB_PATH=$1 timeout 120 $B --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['sanity']['mass_balance_rel_err'], d['sanity']['bad_values'])"; }
for rep in 1 2; do for w in $W; do for l in "$@"; do run $l $w; done; done; done
