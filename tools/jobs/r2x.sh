#!/bin/bash
# round-2 job X (1 GPU): the committed tree once more -- smoke(), all GPU tests
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8) > gpurun_out/r2x_smoke.log
tail -6 gpurun_out/r2x_smoke.log
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2x_tests.log
tail -3 gpurun_out/r2x_tests.log
