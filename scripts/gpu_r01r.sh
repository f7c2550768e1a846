set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r_smoke.log 2>&1; tail -3 gpurun_out/r_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r_pytest.log 2>&1; tail -5 gpurun_out/r_pytest.log
B="timeout 300 python bench.py --n 200 --steps 10 --no-cpu --no-e2e"
for m in neohookean elastic; do for fl in 0 2; do
  $B --flags $fl --material $m > gpurun_out/r_n200_${m}_f${fl}.json 2>&1
done; done
$B --workload twoblock > gpurun_out/r_n200_twoblock.json 2>&1
$B --assembly ordered > gpurun_out/r_n200_neohookean_ordered_f2.json 2>&1
for f in gpurun_out/r_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01r_ner_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 > gpurun_out/ncu_r_f2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01r_elastic_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 --material elastic > gpurun_out/ncu_r_el.log 2>&1
