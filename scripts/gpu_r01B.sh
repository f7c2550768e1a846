set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/B_smoke.log 2>&1; tail -3 gpurun_out/B_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/B_pytest.log 2>&1; tail -4 gpurun_out/B_pytest.log
B="timeout 300 python bench.py --n 200 --steps 20 --no-cpu --no-e2e"
$B --assembly ordered > gpurun_out/B_n200_neohookean_ordered.json 2>&1
$B --assembly ordered --material elastic > gpurun_out/B_n200_elastic_ordered.json 2>&1
$B > gpurun_out/B_n200_neohookean_atomic.json 2>&1
for f in gpurun_out/B_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
