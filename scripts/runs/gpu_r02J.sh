# scripts/gpu_r02J.sh — elastic element kernel as 4 CTAs x 128 threads per SM (ElemShape<0>): parity tests, A/B lines,
# re-stamped ncu traffic (the kernel sources changed), one B200
set -x
T=r02J
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3
for MAT in elastic neohookean; do
  timeout 300 python bench.py --n 200 --material $MAT --steps 20 --no-e2e --no-cpu > gpurun_out/${T}_bench_n200_${MAT}.json 2> gpurun_out/${T}_bench_n200_${MAT}.err
  python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n200_${MAT}.json')); print('LINE $MAT step %.3f ms elem %.3f ms fp64 %.4f parity %s clocks %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['fp64']['frac'], d['parity']['max_rel_f'], d['clocks']))"
done
timeout 600 python bench.py --workload twoblock --steps 10 --no-cpu --no-e2e > gpurun_out/${T}_bench_n400_twoblock.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_n400_twoblock.json')); print('LINE twoblock value %.4e step %.3f ms elem %.3f fp64 %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['fp64']['frac']))"
bash scripts/ncu_traffic.sh ${T} > gpurun_out/${T}_ncu_traffic.log 2>&1; tail -1 gpurun_out/${T}_ncu_traffic.log | cut -c1-200
