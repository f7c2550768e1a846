set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/J_pytest.log 2>&1; tail -6 gpurun_out/J_pytest.log
B="timeout 300 python bench.py --n 200 --steps 10 --no-cpu --no-e2e"
$B --shuffle --flags 6 > gpurun_out/J_n200_neohookean_shuffled_reordered.json 2>&1
$B --shuffle --flags 6 --material elastic > gpurun_out/J_n200_elastic_shuffled_reordered.json 2>&1
$B --flags 6 > gpurun_out/J_n200_neohookean_lattice_reordered.json 2>&1
$B --flags 2 > gpurun_out/J_n200_neohookean_lattice.json 2>&1
for f in gpurun_out/J_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
