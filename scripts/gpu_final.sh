# scripts/gpu_final.sh — the round's closing measurements on one B200: smoke, the GPU test suite, the default bench
# line (64 M neohookean) with the reference arm beside it, the ncu launch list of the same command, one ncu --set full
# capture of the element kernel, and the size-independent properties at the 64 M-element headline size.
set -x
T=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; tail -3 gpurun_out/${T}_bench_n400.err; cut -c1-400 gpurun_out/${T}_bench_n400.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>&1; cut -c1-300 gpurun_out/${T}_bench_ref.json
timeout 600 python bench.py --n 200 --material elastic --steps 20 > gpurun_out/${T}_bench_n200_elastic.json 2>&1; cut -c1-200 gpurun_out/${T}_bench_n200_elastic.json
timeout 600 python bench.py --workload twoblock --steps 10 --no-cpu > gpurun_out/${T}_bench_n400_twoblock.json 2>&1; cut -c1-200 gpurun_out/${T}_bench_n400_twoblock.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_n400.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_${T}_neo_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 > gpurun_out/${T}_ncu_full.log 2>&1
NSM_FULL_SIZE_N=400 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "full_size and neohookean" > gpurun_out/${T}_fullsize_n400.log 2>&1; tail -3 gpurun_out/${T}_fullsize_n400.log
ls -la gpurun_out | tail -14
