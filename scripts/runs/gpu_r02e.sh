# scripts/gpu_r02e.sh — round 2, fifth GPU call: full suite on the frozen kernels, default bench + reference arm,
# configs[1], ncu launch list + stamped ncu traffic capture, compute-sanitizer.
set -x
T=r02e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
( time timeout 1800 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/${T}_pytest.log 2>&1; tail -16 gpurun_out/${T}_pytest.log
bash scripts/ncu_traffic.sh $T; cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json
timeout 900 python bench.py > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; echo rc=$?; tail -3 gpurun_out/${T}_bench_n400.err; cut -c1-400 gpurun_out/${T}_bench_n400.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>&1; cut -c1-300 gpurun_out/${T}_bench_ref.json
bash scripts/bench_config1.sh $T
timeout 900 python bench.py --workload twoblock --steps 10 --no-cpu > gpurun_out/${T}_bench_n400_twoblock.json 2> gpurun_out/${T}_bench_n400_twoblock.err; echo rc=$?; cut -c1-300 gpurun_out/${T}_bench_n400_twoblock.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_n400.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/${T}_ncu_launch.log 2>&1
bash scripts/gpu_sanitize.sh $T
ls -la gpurun_out | tail -30
