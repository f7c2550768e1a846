#!/bin/bash
# after the set_contact change (arguments checked before the entities in place are released, sort scratch accounted):
# the whole GPU suite once more
T=r02Y
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
