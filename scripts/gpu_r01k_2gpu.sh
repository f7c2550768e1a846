set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/k_smoke.log 2>&1; tail -3 gpurun_out/k_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; tail -15 gpurun_out/k_pytest.log
B="timeout 300 python bench.py --n 200 --steps 5 --no-cpu --no-e2e"
$B --flags 2 > gpurun_out/k_n200_neo_f2.json 2>&1
$B --flags 0 > gpurun_out/k_n200_neo_f0.json 2>&1
$B --flags 2 --material elastic > gpurun_out/k_n200_elastic_f2.json 2>&1
for f in gpurun_out/k_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --edge 200 --steps 10 --warmup 3 --no-e2e > gpurun_out/k_bench_2gpu_n200.json 2> gpurun_out/k_bench_2gpu_n200.err; tail -3 gpurun_out/k_bench_2gpu_n200.err; cut -c1-300 gpurun_out/k_bench_2gpu_n200.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01k_neo_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 > gpurun_out/ncu_k_f2.log 2>&1
