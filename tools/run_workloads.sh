#!/bin/bash
# Runs bench.py workloads on the N GPUs of this box and keeps one JSON line per run in gpurun_out/:
#   tools/run_workloads.sh N "c4 c5 c2strong c2" [extra bench.py flags]
N=$1; shift
LIST=$1; shift
EXTRA="$@"
for w in $LIST; do
  case $w in
    c2strong) ARGS="--workload c2 --scaling strong"; TAG=c2_strong ;;
    *) ARGS="--workload $w"; TAG=$w ;;
  esac
  OUT=gpurun_out/r2_${TAG}_n${N}.json
  if [ "$N" -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 $ARGS $EXTRA > $OUT 2> gpurun_out/r2_${TAG}_n${N}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) \
      bench.py --gpus $N --steps 5 --warmup 3 $ARGS $EXTRA > $OUT 2> gpurun_out/r2_${TAG}_n${N}.err
  fi
  echo "== $TAG N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT").read().strip().splitlines()[-1])
    print("  value %.0f %s, %.2f ms/step, e2e %s, parity %s, collectives %s" % (d["value"], d["unit"], d["ms_per_step"],
          d.get("e2e", {}).get("value"), d.get("parity"), d.get("run", {}).get("collective_ms_per_step")))
except Exception as e:
    print("  no JSON line:", e)
    print(open("gpurun_out/r2_${TAG}_n${N}.err").read()[-1500:])
PY
done
