# scripts/gpu_r02a.sh — round 2, first GPU call: the GPU suite (new: ATOMIC decks, headline-size sampled parity,
# multi-rank tests in lockstep on one GPU), the issue-model microbenchmark, the default bench line with its parity
# block, the reference arm, and the stamped ncu traffic capture.
set -x
T=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
( time timeout 1800 python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/${T}_pytest.log 2>&1; tail -30 gpurun_out/${T}_pytest.log
timeout 300 ./scripts/micro/issue_mix > gpurun_out/${T}_issue_mix.txt 2>&1; cat gpurun_out/${T}_issue_mix.txt
timeout 900 python bench.py > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; echo rc=$?; tail -3 gpurun_out/${T}_bench_n400.err; cut -c1-600 gpurun_out/${T}_bench_n400.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>&1; cut -c1-600 gpurun_out/${T}_bench_ref.json
bash scripts/ncu_traffic.sh $T; cat gpurun_out/ncu_traffic.json | head -30
ls -la gpurun_out | tail -20
