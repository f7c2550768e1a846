set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -10; free -g | head -2; nproc
timeout 600 python -m pytest tests/test_gpu_host_cpp.py -m gpu -q -x -k decomposed > gpurun_out/s_pytest.log 2>&1; tail -3 gpurun_out/s_pytest.log
run() { # nproc edge steps tag extra
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --edge $2 --steps $3 --warmup 3 --no-e2e $5 > gpurun_out/s_$4.json 2> gpurun_out/s_$4.err; tail -2 gpurun_out/s_$4.err
}
run 8 200 20 8gpu_n200_neohookean_f2
run 4 200 20 4gpu_n200_neohookean_f2
run 8 200 20 8gpu_n200_twoblock "--workload twoblock"
run 8 400 10 8gpu_n400_neohookean_f2
timeout 300 python bench.py --n 200 --steps 20 --no-cpu --no-e2e > gpurun_out/s_1gpu_n200_neohookean_f2.json 2>&1
for f in gpurun_out/s_*gpu_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  ms/step %.3f elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
