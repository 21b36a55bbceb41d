#!/bin/bash
# usage: tools/multi_gpu_run.sh <N> : C3 (full line, peer delivery), C3 with the NCCL gather, C4vg peer / nccl, on N GPUs of one box
N=$1
run() { # name, extra args
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/r2_${name}_${N}gpu.json 2> gpurun_out/r2_${name}_${N}gpu.err
  python - <<PY
import json
try:
    j=[json.loads(l) for l in open("gpurun_out/r2_${name}_${N}gpu.json") if l.startswith("{")][-1]
    print("${name} N=$N", round(j["value"],1), "Mrays/s  ms/step", round(j["ms_per_step"],2), "kernel_ms(max)", round(j["roofline"]["kernel_ms_max_over_ranks"],2), "e2e", (j.get("e2e") or {}).get("value"), j.get("parity"), "|", j.get("result_delivery"))
except Exception as e:
    print("${name} N=$N failed", e); print(open("gpurun_out/r2_${name}_${N}gpu.err").read()[-1500:])
PY
}
run c3_peer --steps 5 --warmup 3 --cpu-seconds 6
run c3_nccl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gather nccl
run c4vg_peer --config C4vg --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
run c4vg_nccl --config C4vg --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gather nccl
run c2_weak --config C2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
