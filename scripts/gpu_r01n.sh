set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/n_smoke.log 2>&1; tail -3 gpurun_out/n_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/n_pytest.log 2>&1; tail -5 gpurun_out/n_pytest.log
B="timeout 300 python bench.py --n 200 --steps 10 --no-cpu --no-e2e"
for m in neohookean elastic; do
  $B --flags 2 --material $m > gpurun_out/n_n200_${m}_f2.json 2>&1
  for v in t256b2_nostage t192b3_nostage t192b3_r112_nostage t160b4_nostage; do
    NSM_B200_LIB=$PWD/nimblesm_b200/lib/variants/libnsm_b200_$v.so $B --flags 2 --material $m > gpurun_out/n_n200_${m}_f2_$v.json 2>&1
  done
done
$B --flags 0 --material neohookean > gpurun_out/n_n200_neohookean_f0.json 2>&1
$B --flags 0 --material elastic > gpurun_out/n_n200_elastic_f0.json 2>&1
$B --workload twoblock > gpurun_out/n_n200_twoblock.json 2>&1
for f in gpurun_out/n_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
