# scripts/gpu_r02c.sh — round 2, third GPU call: the whole GPU suite after the state-slot / ENTRYCONST / host-barrier /
# reference-binding work.
set -x
T=r02c
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${T}_pytest.log 2>&1; tail -40 gpurun_out/${T}_pytest.log
