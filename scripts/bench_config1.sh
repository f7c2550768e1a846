# scripts/bench_config1.sh — BASELINE.json configs[1]: synthetic structured hex8 cube, 8 M elements, linear elastic,
# 1000 explicit steps on 1 B200.  One clock-sampled bench line (bench.py exits non-zero when no nvidia-smi sample
# fell inside the timed region or the parity block fails) -> profiles/<tag>_bench_config1_8M_elastic_1000steps.json
T=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python bench.py --material elastic --n 200 --steps 1000 --warmup 3 > gpurun_out/${T}_bench_config1_8M_elastic_1000steps.json \
  2> gpurun_out/${T}_bench_config1.err
echo "config1 rc=$?"; tail -2 gpurun_out/${T}_bench_config1.err; cut -c1-300 gpurun_out/${T}_bench_config1_8M_elastic_1000steps.json
