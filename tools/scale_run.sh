#!/bin/bash
# Multi-GPU bench lines of one box (profiles/r2_*): usage tools/scale_run.sh N [rmat_scale ...]
# default weak-scaling line, BASELINE configs[4] (200^3, strong), configs[3] (R-MAT, strong) at the given scales
N=$1; shift
SCALES=${@:-22}
mkdir -p gpurun_out
run() {  # name, args...
  name=$1; shift
  if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "== $name rc=$? $(tail -c 300 gpurun_out/$name.err | tr '\n' ' ')"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/$name.json").read().strip().splitlines()[-1])
    print("   ", d["n_gpus"], "GPUs", round(d["ms_per_step"],3), "ms/step", round(d["value"]/1e9,2), "G products/s", "frac", round(d["roofline"]["frac"],4), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), [ (p["rank"], round(p["ms_mean"],2)) for p in d.get("per_rank",[])][:8])
except Exception as e: print("    no line:", e)
P
}
run r2_weak30_n$N --steps 10 --warmup 3 --no-cpu-baseline
run r2_strong200_n$N --strong --side 200 --max-power 5 --no-e2e --no-cpu-baseline --steps 5 --warmup 4
for s in $SCALES; do run r2_rmat${s}_n$N --workload rmat --scale $s --steps 5 --warmup 3; done
