set -x
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_n400.json 2> gpurun_out/bench_n400.err; tail -3 gpurun_out/bench_n400.err; cat gpurun_out/bench_n400.json
for v in sd mb3 sd_mb3; do NSM_B200_LIB=$PWD/nimblesm_b200/lib/variants/libnsm_b200_$v.so timeout 200 python bench.py --n 200 --steps 5 --no-cpu --no-e2e > gpurun_out/bench_n200_$v.json 2>&1; done
timeout 200 python bench.py --n 200 --steps 5 --no-cpu --no-e2e --flags 2 > gpurun_out/bench_n200_binv.json 2>&1
timeout 200 python bench.py --n 200 --steps 5 --no-cpu --no-e2e --assembly ordered > gpurun_out/bench_n200_ordered.json 2>&1
timeout 200 python bench.py --n 200 --steps 5 --no-cpu --no-e2e --material elastic > gpurun_out/bench_n200_elastic.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f -o gpurun_out/prof_elem_r01 python bench.py --n 200 --steps 1 --no-e2e --no-cpu > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out
