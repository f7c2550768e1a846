# scripts/bench_config4.sh [tag] [gpus] [edge] — BASELINE.json configs[4]: multi-block heterogeneous-material mesh
# (elastic + neohookean blocks per GPU) with prescribed-velocity BCs on both x faces, weak scaling over the GPUs of one
# box.  One clock-sampled line with the parity block (replicas bit-equal across ranks + a window straddling the
# partition faces AND the material interface vs the oracle) -> profiles/<tag>_bench_config4_<gpus>gpu_twoblock.json
T=${1:-r02}; N=${2:-8}; E=${3:-400}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --workload twoblock --edge $E --steps 10 --warmup 3 --no-cpu \
  > gpurun_out/${T}_bench_config4_${N}gpu_twoblock.json 2> gpurun_out/${T}_bench_config4_${N}gpu.err
echo "config4 rc=$?"; tail -3 gpurun_out/${T}_bench_config4_${N}gpu.err; cut -c1-300 gpurun_out/${T}_bench_config4_${N}gpu_twoblock.json
