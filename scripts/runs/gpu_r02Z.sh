#!/bin/bash
# the new contact_entity_creation driver test, the set_contact argument test, one reference deck through the driver
# (host material factory changed: string parameters, first value of a repeated key wins)
T=r02Z
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_contact.py tests/test_gpu_host_cpp.py -m gpu -q -k "entity_creation or argument_errors or (driver_runs_reference_decks and brick_with_fibers)" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${T}_pytest.log
