set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/m_smoke.log 2>&1; tail -3 gpurun_out/m_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/m_pytest.log 2>&1; tail -15 gpurun_out/m_pytest.log
timeout 900 python bench.py > gpurun_out/m_bench_n400.json 2> gpurun_out/m_bench_n400.err; tail -3 gpurun_out/m_bench_n400.err; cut -c1-600 gpurun_out/m_bench_n400.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/m_bench_ref.json 2>&1; cut -c1-400 gpurun_out/m_bench_ref.json
timeout 600 python bench.py --n 200 --material elastic --steps 20 > gpurun_out/m_bench_n200_elastic.json 2>&1; cut -c1-300 gpurun_out/m_bench_n200_elastic.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/m_launches_n400.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/m_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01m_neo_f2 python bench.py --n 200 --steps 1 --no-e2e --no-cpu --flags 2 > gpurun_out/ncu_m_f2.log 2>&1
ls -la gpurun_out | tail -12
