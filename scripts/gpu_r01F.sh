set -x
mkdir -p gpurun_out
B="timeout 300 python bench.py --n 200 --steps 20 --no-cpu --no-e2e"
for m in elastic neohookean; do
  $B --flags 2 --material $m > gpurun_out/F_n200_${m}_f2.json 2>&1
  NSM_B200_LIB=$PWD/nimblesm_b200/lib/variants/libnsm_b200_pretrim.so $B --flags 2 --material $m > gpurun_out/F_n200_${m}_f2_pretrim.json 2>&1
  $B --flags 2 --material $m > gpurun_out/F_n200_${m}_f2_again.json 2>&1
done
for f in gpurun_out/F_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
