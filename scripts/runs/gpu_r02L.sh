# scripts/gpu_r02L.sh — last check of HEAD on one B200: smoke, GPU suite, default bench line + reference arm, configs[1]
set -x
T=r02L
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -4 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; tail -3 gpurun_out/${T}_bench_n400.err; cut -c1-260 gpurun_out/${T}_bench_n400.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>&1; cut -c1-200 gpurun_out/${T}_bench_ref.json
bash scripts/bench_config1.sh ${T}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_n400.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
