#!/bin/bash
# 1/2/4/8-GPU bench lines on ONE 8-GPU box (strong scaling of the same 10^6-parameter training set), launched exactly as
# the driver does.  Outputs gpurun_out/<tag>_bench_n{1,2,4,8}.json.   usage (on the box): bash tools/scaling_run.sh <tag>
TAG=${1:-scale}
OUT=gpurun_out
mkdir -p $OUT
timeout -s KILL 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
tail -c 400 $OUT/${TAG}_bench_n1.json; echo
for N in 2 4 8; do
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.0f e2e %.0f ms/step %.1f argmax-ok" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print("N=$N failed", e)
PY
done
