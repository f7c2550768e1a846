set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/C_pytest_2gpu.log 2>&1; tail -4 gpurun_out/C_pytest_2gpu.log
T="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
$T --edge 200 --steps 20 --warmup 3 --no-e2e > gpurun_out/C_2gpu_n200_neohookean.json 2> gpurun_out/C_2gpu_a.err; tail -2 gpurun_out/C_2gpu_a.err
$T --edge 200 --steps 20 --warmup 3 --no-e2e --assembly ordered > gpurun_out/C_2gpu_n200_neohookean_ordered.json 2> gpurun_out/C_2gpu_b.err; tail -2 gpurun_out/C_2gpu_b.err
$T --steps 5 --warmup 3 > gpurun_out/C_2gpu_n400_default.json 2> gpurun_out/C_2gpu_c.err; tail -2 gpurun_out/C_2gpu_c.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/C_2gpu_ref.json 2> gpurun_out/C_2gpu_d.err; cut -c1-200 gpurun_out/C_2gpu_ref.json
for f in gpurun_out/C_2gpu_n*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  ms/step %.3f elem_ms %.3f node_ms %.3f e2e %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("e2e",{}).get("value")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
