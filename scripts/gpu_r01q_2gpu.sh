set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -8
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; tail -5 gpurun_out/q_pytest.log
B="timeout 300 python bench.py --n 200 --steps 10 --no-cpu --no-e2e"
$B --flags 2 > gpurun_out/q_n200_neohookean_f2.json 2>&1
T="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-e2e"
$T --edge 200 --steps 20 --warmup 3 > gpurun_out/q_2gpu_n200_neohookean_f2.json 2> gpurun_out/q_2gpu_n200.err; tail -3 gpurun_out/q_2gpu_n200.err
$T --edge 200 --steps 20 --warmup 3 --material elastic > gpurun_out/q_2gpu_n200_elastic_f2.json 2> gpurun_out/q_2gpu_n200e.err; tail -3 gpurun_out/q_2gpu_n200e.err
$T --edge 200 --steps 20 --warmup 3 --workload twoblock > gpurun_out/q_2gpu_n200_twoblock.json 2> gpurun_out/q_2gpu_n200t.err; tail -3 gpurun_out/q_2gpu_n200t.err
$T --edge 400 --steps 10 --warmup 3 > gpurun_out/q_2gpu_n400_neohookean_f2.json 2> gpurun_out/q_2gpu_n400.err; tail -3 gpurun_out/q_2gpu_n400.err
for f in gpurun_out/q_*n?00_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  ms/step %.3f elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
for m in neohookean elastic; do for fl in 0 2; do
timeout 300 ncu --metrics sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:element_force -s 3 -c 1 --csv --log-file gpurun_out/q_dp_${m}_f${fl}.csv python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags $fl --material $m > /dev/null 2>&1
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01q_elastic_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 --material elastic > gpurun_out/ncu_q_el.log 2>&1
