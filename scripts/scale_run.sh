#!/bin/bash
# the driver's scaling sweep, as it launches it: usage: gpurun --gpus 8 -- bash scripts/scale_run.sh 8 4
out=gpurun_out/r02; mkdir -p $out
for n in "$@"; do
  t0=$(date +%s); python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
     --master-port $((29540 + n)) bench.py --gpus $n --steps 20 --warmup 5 > $out/scale_n$n.json 2> $out/scale_n$n.err
  echo "N=$n wall $(( $(date +%s) - t0 )) s"; tail -2 $out/scale_n$n.err | cut -c1-300
  python - <<EOF
import json
ls=[l for l in open("$out/scale_n$n.json") if l.startswith("{")]
if ls:
    d=json.loads(ls[-1])
    print("N=$n", round(d["value"],2), [round(x,4) for x in d["ms_per_step_by_rank"]], "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],4), d["sanity"].get("slab_equals_single"), d["sanity"].get("bad_values"))
    for k in ("config3","config4"):
        b=d.get(k); print(" ", k, b and (round(b["value"],1), round(b["ms_per_step"],3), b.get("roofline_frac"), b["sanity"]["mass_balance_rel_err"], b["sanity"]["bad_values"]))
else:
    print("N=$n: no JSON line")
EOF
done
