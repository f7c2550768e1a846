# scripts/gpu_r02q.sh — pipelined host paths: CTA slots left to the copy-side kernels (NSM_B200_PIPE_RESERVE sweep), one B200
set -x
T=${1:-r02q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pipelined or step_host" 2>&1 | tail -2
for R in 0 8 16 32 64; do
  NSM_B200_PIPE_RESERVE=$R timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/${T}_bench_n400_reserve${R}.json 2> gpurun_out/${T}_bench_n400_reserve${R}.err
  python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n400_reserve${R}.json')); e=d['e2e']; print('RESERVE $R value %.4e e2e %.4e (%.1f ms) seam %.4e (%.1f ms)' % (d['value'], e['value'], e['ms_per_step'], e['force_seam']['value'], e['force_seam']['ms_per_call']))"
done
