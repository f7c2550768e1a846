set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/g_smoke.log 2>&1; tail -4 gpurun_out/d_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; tail -15 gpurun_out/d_pytest.log
B="timeout 200 python bench.py --n 200 --steps 5 --no-cpu --no-e2e"
$B --flags 0 > gpurun_out/g_n200_default_f0.json 2>&1
$B --flags 2 > gpurun_out/g_n200_default_f2.json 2>&1
for v in t128b4 t384b1; do
  NSM_B200_LIB=$PWD/nimblesm_b200/lib/variants/libnsm_b200_$v.so $B --flags 2 > gpurun_out/g_n200_${v}_f2.json 2>&1
done
$B --flags 2 --material elastic > gpurun_out/g_n200_elastic_f2.json 2>&1
$B --flags 2 --assembly ordered > gpurun_out/g_n200_ordered_f2.json 2>&1
for f in gpurun_out/g_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01g_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 > gpurun_out/ncu_g_f2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01g_elastic python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 --material elastic > gpurun_out/ncu_g_el.log 2>&1
ls -la gpurun_out | tail -5
