#!/bin/bash
# usage: tools/multi_gpu_run.sh <N> [names...] : bench lines on N GPUs of one box (default: c3_peer c3_nccl c4vg_peer c4vg_nccl c2_weak)
N=$1; shift
names=${@:-c3_peer c3_nccl c4vg_peer c4vg_nccl c2_weak}
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
for n in $names; do
  case $n in
    c3_driver) run c3_driver --steps 20 --warmup 5 ;;
    c3_peer) run c3_peer --steps 5 --warmup 3 --cpu-seconds 6 ;;
    c3_peer_fast) run c3_peer_fast --steps 10 --warmup 3 --no-cpu-baseline --no-e2e ;;
    c3_store) run c3_store --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gather peer-store ;;
    c3_nccl) run c3_nccl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gather nccl ;;
    c4vg_peer) run c4vg_peer --config C4vg --steps 5 --warmup 3 --no-cpu-baseline --no-e2e ;;
    c4vg_nccl) run c4vg_nccl --config C4vg --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --gather nccl ;;
    c4kd_peer) run c4kd_peer --config C4kd --steps 5 --warmup 3 --no-cpu-baseline --no-e2e ;;
    c2_weak) run c2_weak --config C2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e ;;
  esac
done
