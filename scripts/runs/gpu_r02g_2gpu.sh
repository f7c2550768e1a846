# scripts/gpu_r02g_2gpu.sh — round 2, two GPUs: the multi-rank tests on real NVLink peers, and the torchrun bench lines
# (cube + two-block) with their parity blocks (replicas bit-equal, window straddling the partition face vs the oracle).
set -x
T=r02g
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_host_cpp.py -m gpu -q -k "decomposed or decomposes" ) > gpurun_out/${T}_pytest_2gpu.log 2>&1; tail -8 gpurun_out/${T}_pytest_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_2gpu_n400.json 2> gpurun_out/${T}_bench_2gpu_n400.err
echo rc=$?; tail -3 gpurun_out/${T}_bench_2gpu_n400.err; cut -c1-300 gpurun_out/${T}_bench_2gpu_n400.json
bash scripts/bench_config4.sh $T 2 400
python - <<'PY'
import json
for f in ("r02g_bench_2gpu_n400", "r02g_bench_config4_2gpu_twoblock"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.4g ms %.3f clocks %s\n   e2e %s\n   parity %s" % (d["value"], d["ms_per_step"], d["clocks"], {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_gbs_per_rank", "d2h_gbs_per_rank", "numa_binding")}, d.get("parity")))
    except Exception as ex:
        print(f, "failed", ex)
PY
