#!/bin/bash
# Final check of the committed state, as the driver runs it: strict GPU suite, smoke(), default bench line, reference arm.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err; cut -c1-260 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-260
