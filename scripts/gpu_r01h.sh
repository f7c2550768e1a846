set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/h_smoke.log 2>&1; tail -3 gpurun_out/h_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; tail -5 gpurun_out/h_pytest.log
B="timeout 300 python bench.py --n 200 --steps 5 --no-cpu --no-e2e"
$B --flags 0 > gpurun_out/h_n200_neo_f0.json 2>&1
$B --flags 2 > gpurun_out/h_n200_neo_f2.json 2>&1
$B --flags 2 --material elastic > gpurun_out/h_n200_elastic_f2.json 2>&1
$B --flags 0 --material elastic > gpurun_out/h_n200_elastic_f0.json 2>&1
$B --flags 2 --assembly ordered > gpurun_out/h_n200_neo_ordered_f2.json 2>&1
timeout 900 python bench.py > gpurun_out/h_bench_n400.json 2> gpurun_out/h_bench_n400.err; tail -3 gpurun_out/h_bench_n400.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_reference.json 2> gpurun_out/h_bench_reference.err
for f in gpurun_out/h_n200_*.json gpurun_out/h_bench_n400.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s e2e %s cpu %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points"),d.get("e2e",{}).get("value"),d.get("cpu_baseline",{}).get("value")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
cat gpurun_out/h_bench_reference.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01h_launches_bench_n400.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_h_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01h_neo_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 > gpurun_out/ncu_h_f2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01h_neo_f0 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 0 > gpurun_out/ncu_h_f0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01h_elastic_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 --material elastic > gpurun_out/ncu_h_el.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:node_ -s 6 -c 2 -f -o gpurun_out/prof_node_r01h python bench.py --n 200 --steps 4 --no-e2e --no-cpu > gpurun_out/ncu_h_node.log 2>&1
ls -la gpurun_out | tail -8
