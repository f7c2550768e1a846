# scripts/gpu_r02i_8gpu.sh — round 2, eight GPUs: BASELINE configs[3] (512 M-element neohookean cube) and configs[4]
# (two-block elastic + neohookean, prescribed velocity on both x faces) under torchrun, each with clocks sampled in the
# timed region, the parity block (replicas + straddling window vs oracle) and the host<->device copy ceiling of the box.
set -x
T=r02i
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 8 --steps 10 --warmup 3 --copy-ceiling > gpurun_out/${T}_bench_8gpu_n400.json 2> gpurun_out/${T}_bench_8gpu_n400.err
echo rc=$?; tail -3 gpurun_out/${T}_bench_8gpu_n400.err
bash scripts/bench_config4.sh $T 8 400
python - <<'PY'
import json
for f in ("r02i_bench_8gpu_n400", "r02i_bench_config4_8gpu_twoblock"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.3f clocks %s" % (d["value"], d["ms_per_step"], d["clocks"]))
        print("   e2e", {k: d["e2e"].get(k) for k in ("value", "ms_per_step", "h2d_gbs_per_rank", "d2h_gbs_per_rank", "host_traffic_gbs_all_ranks", "numa_binding", "copy_ceiling")})
        print("   parity", d.get("parity"))
    except Exception as ex:
        print(f, "failed", ex)
PY
