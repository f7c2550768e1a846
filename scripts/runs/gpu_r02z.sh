# scripts/gpu_r02z.sh — pipelined host step: node-chunk sweep with the copy-side CTA reservation in place, one B200
set -x
mkdir -p gpurun_out
for C in 8 16 24 32 48; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-parity --host-chunks $C > gpurun_out/r02z_bench_n400_chunks${C}.json 2> gpurun_out/r02z_bench_n400_chunks${C}.err
  python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_n400_chunks${C}.json')); e=d['e2e']; print('CHUNKS $C e2e %.4e (%.1f ms) up %.1f down %.1f GB/s; seam %.4e (%.1f ms)' % (e['value'], e['ms_per_step'], e['h2d_gbs_per_rank'], e['d2h_gbs_per_rank'], e['force_seam']['value'], e['force_seam']['ms_per_call']))"
done
