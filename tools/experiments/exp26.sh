#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
F2B_PARITY_LOG=$PWD/gpurun_out/parity.log timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_parity.log; wc -l gpurun_out/parity.log
