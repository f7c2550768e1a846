set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
B="timeout 200 python bench.py --n 200 --steps 5 --no-cpu --no-e2e"
$B > gpurun_out/b_n200_default.json 2>&1
$B --flags 2 > gpurun_out/b_n200_default_binv.json 2>&1
for v in plaindiv t256b3 t128b5 t128b6 t128b4 t64b10; do
  NSM_B200_LIB=$PWD/nimblesm_b200/lib/variants/libnsm_b200_$v.so $B > gpurun_out/b_n200_$v.json 2>&1
  NSM_B200_LIB=$PWD/nimblesm_b200/lib/variants/libnsm_b200_$v.so $B --flags 2 > gpurun_out/b_n200_${v}_binv.json 2>&1
done
$B --material elastic > gpurun_out/b_n200_elastic.json 2>&1
$B --assembly ordered > gpurun_out/b_n200_ordered.json 2>&1
for f in gpurun_out/b_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"]))
except Exception as e: print("ERR",e, open("$f").read()[-500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01b python bench.py --n 200 --steps 1 --no-e2e --no-cpu > gpurun_out/ncu_full_run_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01b_binv python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 > gpurun_out/ncu_full_run_b2.log 2>&1
ls -la gpurun_out
