# same-box A/B of kernel variants built with scripts/build_variant.sh (default library first)
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-parity --no-extra"
for lib in "" $(ls topoflow36_b200/_variants/*.so 2>/dev/null); do
  export TFB_LIB_PATH=$lib; [ -z "$lib" ] && unset TFB_LIB_PATH
  for w in ${AB_WORKLOADS:-kw kw_ga dw dw_full}; do
    echo -n "lib=${lib##*/} $w: "; timeout 120 $B --workload $w 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms_per_step'], round(d['roofline']['frac'],4), d['sanity']['mass_balance_rel_err'], d['sanity']['bad_values'])"
  done
done
