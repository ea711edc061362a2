#!/bin/bash
# multi-GPU scaling of bench.py on ONE box: usage r02_scaling.sh "<list of N>"  (gpurun --gpus max(N)).  Weak: 4096 instances per GPU;
# strong: 4096 (and 32768) instances in total.  One process per GPU under torchrun, exactly the driver's launch line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
port=29611
for n in $1; do
  for mode in ${MODES:-weak strong strong32k}; do
    extra="--scaling weak"
    [ $mode == strong ] && extra="--scaling strong"
    [ $mode == strong32k ] && extra="--scaling strong --batch 32768"
    out=gpurun_out/r02s_${mode}_n${n}.json
    if [ $n == 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline $extra > $out 2> ${out%.json}.err
    else
      port=$((port+1))
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline $extra > $out 2> ${out%.json}.err
    fi
    python - $out $mode $n <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-10s N=%s value %.4g ms/step %.4g e2e %.4g batch %s" % (sys.argv[2], sys.argv[3], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d["config"].get("global_batch")))
except Exception as e: print("parse failed", sys.argv[1], e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
  done
done
