# scripts/gpu_r02s.sh — checkpoint on one B200 after the contact / force-seam / host-pipeline work: smoke, GPU suite,
# stamped ncu traffic of the element kernels (the node-kernel argument block changed, so the source hash did), the
# default bench line with the reference arm beside it, the ORDERED line, configs[1], the launch list.
set -x
T=${1:-r02s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
bash scripts/ncu_traffic.sh ${T} > gpurun_out/${T}_ncu_traffic.log 2>&1; tail -2 gpurun_out/${T}_ncu_traffic.log | cut -c1-400
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json
timeout 900 python bench.py > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; tail -3 gpurun_out/${T}_bench_n400.err; cut -c1-300 gpurun_out/${T}_bench_n400.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>&1; cut -c1-300 gpurun_out/${T}_bench_ref.json
timeout 600 python bench.py --assembly ordered --steps 10 --no-cpu > gpurun_out/${T}_bench_n400_ordered.json 2>&1; cut -c1-200 gpurun_out/${T}_bench_n400_ordered.json
timeout 600 python bench.py --workload twoblock --steps 10 --no-cpu > gpurun_out/${T}_bench_n400_twoblock.json 2>&1; cut -c1-200 gpurun_out/${T}_bench_n400_twoblock.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_n400.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
ls -la gpurun_out | tail -5
