set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_host_cpp.py -q -k "decomposed" > gpurun_out/j_pytest_2gpu.log 2>&1; tail -15 gpurun_out/j_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --n 200 --steps 10 --warmup 3 --no-e2e > gpurun_out/j_bench_2gpu_n200.json 2> gpurun_out/j_bench_2gpu_n200.err; tail -5 gpurun_out/j_bench_2gpu_n200.err; cat gpurun_out/j_bench_2gpu_n200.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/j_bench_2gpu_n400.json 2> gpurun_out/j_bench_2gpu_n400.err; tail -5 gpurun_out/j_bench_2gpu_n400.err; cat gpurun_out/j_bench_2gpu_n400.json | cut -c1-400
timeout 600 python bench.py --gpus 1 --n 200 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/j_bench_1gpu_n200.json 2>&1; cut -c1-200 gpurun_out/j_bench_1gpu_n200.json
