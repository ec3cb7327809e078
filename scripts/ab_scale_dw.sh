#!/bin/bash
# Same-box A/B of the DIFFUSIVE step at N = 1 and N = 2: NCCL halo (between and after the passes)
# against the peer-memory halo fused into both passes.  usage: gpurun --gpus 2 -- bash scripts/ab_scale_dw.sh
out=gpurun_out/r02/ab_scale_dw.log
mkdir -p gpurun_out/r02
: > $out
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], d["ms_per_step_by_rank"], round(d["value"],2), d["sanity"].get("slab_equals_single"))'
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
FL="--no-e2e --no-extra --no-cpu-baseline"
for w in dw dw_full; do
  python bench.py --steps 200 --warmup 20 $FL --no-parity --workload $w 2>>$out.err | python -c "$pick" "N1 $w k200" >> $out
  $TR bench.py --gpus 2 --steps 200 --warmup 20 $FL --workload $w --halo nccl 2>>$out.err | python -c "$pick" "N2 $w nccl k200" >> $out
  $TR bench.py --gpus 2 --steps 200 --warmup 20 $FL --workload $w --halo p2p-dw 2>>$out.err | python -c "$pick" "N2 $w p2p k200" >> $out
  $TR bench.py --gpus 2 --steps 20 --warmup 3 $FL --no-parity --workload $w --halo nccl 2>>$out.err | python -c "$pick" "N2 $w nccl k20" >> $out
  $TR bench.py --gpus 2 --steps 20 --warmup 3 $FL --no-parity --workload $w --halo p2p-dw 2>>$out.err | python -c "$pick" "N2 $w p2p k20" >> $out
done
cat $out; tail -5 $out.err
