#!/bin/bash
# Same-box A/B of the step time at N = 1 and N = 2 (peer-memory halo, NCCL halo), long and
# driver-length (20-step) timed regions.  usage: gpurun --gpus 2 -- bash scripts/ab_scale.sh [reps]
out=gpurun_out/r02/ab_scale.log
mkdir -p gpurun_out/r02
: > $out
pick='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], d["ms_per_step_by_rank"], round(d["value"],2))'
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
FL="--no-e2e --no-parity --no-extra --no-cpu-baseline"
for i in $(seq 1 ${1:-2}); do
  python bench.py --steps 200 --warmup 20 $FL 2>>$out.err | python -c "$pick" "N1 k200" >> $out
  python bench.py --steps 20 --warmup 3 $FL 2>>$out.err | python -c "$pick" "N1 k20" >> $out
  $TR bench.py --gpus 2 --steps 200 --warmup 20 $FL 2>>$out.err | python -c "$pick" "N2 p2p k200" >> $out
  $TR bench.py --gpus 2 --steps 20 --warmup 3 $FL 2>>$out.err | python -c "$pick" "N2 p2p k20" >> $out
  $TR bench.py --gpus 2 --steps 200 --warmup 20 $FL --halo nccl 2>>$out.err | python -c "$pick" "N2 nccl k200" >> $out
done
cat $out; tail -5 $out.err
