#!/bin/bash
# contact visualisation database through the driver + nsm_b200_contact_status against the oracle's flags
T=r02a2
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_contact.py -m gpu -q -k "entity_creation or force_vs_oracle" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/${T}_pytest.log | cut -c1-300
