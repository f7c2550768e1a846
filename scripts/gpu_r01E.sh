set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/E_smoke.log 2>&1; tail -3 gpurun_out/E_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/E_pytest.log 2>&1; tail -4 gpurun_out/E_pytest.log
B="timeout 300 python bench.py --n 200 --steps 20 --no-cpu --no-e2e"
for m in neohookean elastic; do for fl in 2 0; do
  $B --flags $fl --material $m > gpurun_out/E_n200_${m}_f${fl}.json 2>&1
done; done
for f in gpurun_out/E_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
for m in neohookean elastic; do for fl in 0 2; do
timeout 300 ncu --metrics sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:element_force -s 3 -c 1 --csv --log-file gpurun_out/E_dp_${m}_f${fl}.csv python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags $fl --material $m > /dev/null 2>&1
done; done
