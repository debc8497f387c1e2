#!/bin/bash
# usage: tools/scale_sweep.sh N TAG [configs...]   -- bench.py for every config at N GPUs of this box -> gpurun_out/TAG_scale_<cfg>_<N>gpu.json
N=$1; TAG=$2; shift 2
CFGS=${@:-"C2 C3 C4 C5:100000"}
for c in $CFGS; do
  cfg=${c%%:*}; k=${c#*:}; extra=""; name=$cfg
  if [ "$cfg" = "C5" ]; then extra="--c5-k $k"; name="C5_${k}"; fi
  out=gpurun_out/${TAG}_scale_${name}_${N}gpu.json
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 200 --warmup 10 --config $cfg $extra > $out 2> gpurun_out/${TAG}_scale_${name}_${N}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --config $cfg $extra > $out 2> gpurun_out/${TAG}_scale_${name}_${N}gpu.err
  fi
  python - "$out" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("final_gather", {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
