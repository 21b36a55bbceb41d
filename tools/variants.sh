#!/bin/bash
# usage: tools/variants.sh <config> <rays> A B C ...   : bench each hare_b200/libhare_var_<X>.so (tuning experiments; not shipped)
cfg=$1; rays=$2; shift 2
for v in "$@"; do
  HARE_B200_LIB=$PWD/hare_b200/libhare_var_$v.so python bench.py --config $cfg --no-extras --no-cpu-baseline --no-e2e --steps 3 --warmup 2 --rays $rays > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/var_$v.json")); print("variant $v", round(j["value"],1), "Mrays/s kernel_ms", round(j["roofline"]["kernel_ms"],2))
except Exception as e:
    print("variant $v failed", e); print(open("gpurun_out/var_$v.err").read()[-400:])
PY
done
