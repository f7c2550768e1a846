set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "reordered or step_host or multi_step or explicit_steps or bc_programs" > gpurun_out/K_pytest.log 2>&1; tail -8 gpurun_out/K_pytest.log
B="timeout 300 python bench.py --n 200 --steps 10 --no-cpu --no-e2e"
$B --shuffle --flags 14 > gpurun_out/K_n200_neohookean_shuffled_f14.json 2>&1
$B --shuffle --flags 14 --material elastic > gpurun_out/K_n200_elastic_shuffled_f14.json 2>&1
for f in gpurun_out/K_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
