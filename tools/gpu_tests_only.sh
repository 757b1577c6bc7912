#!/bin/bash
tag=${1:-r02r}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
