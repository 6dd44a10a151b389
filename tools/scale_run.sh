#!/bin/bash
# Scaling run on one node: bench.py at N = 1, 2, 4, 8 (as many as the box has), one JSON line each.
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  if [ $n -gt $NG ]; then break; fi
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
d=json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
print("N=%d value=%.0f pairs/s ms_per_step=%.2f e2e=%.0f clocks=%s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
PY
done
