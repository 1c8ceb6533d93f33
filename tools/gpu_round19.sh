#!/bin/bash
mkdir -p gpurun_out
echo "== pytest new"; timeout 400 python -m pytest tests -x -q -m gpu -k "nuclear_cusp or vjp or energy_gradient or paulinet_default or cta_pair" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_new.log | cut -c1-300
