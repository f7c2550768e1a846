# scripts/gpu_r02G_8gpu.sh — closing pass on EIGHT GPUs with the final library: the multi-rank tests on real NVLink peers
# (np2 / np4 decks, in-driver decomposition, contact across partitions) and BASELINE configs[3] (512 M-element
# neohookean cube) under torchrun with clocks, parity block and end-to-end figure.
set -x
T=r02G
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_host_cpp.py tests/test_gpu_contact.py -m gpu -q -k "decomposed or decomposes or across_partitions" ) > gpurun_out/${T}_pytest_multigpu.log 2>&1; tail -6 gpurun_out/${T}_pytest_multigpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_bench_8gpu_n400.json 2> gpurun_out/${T}_bench_8gpu_n400.err
echo rc=$?; tail -3 gpurun_out/${T}_bench_8gpu_n400.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02G_bench_8gpu_n400.json").read().splitlines() if l.startswith("{")][-1])
print("value %.4g ms %.3f clocks %s" % (d["value"], d["ms_per_step"], d["clocks"]))
print("   e2e", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "h2d_gbs_per_rank", "d2h_gbs_per_rank", "host_traffic_gbs_all_ranks")})
print("   parity", d.get("parity"))
PY
