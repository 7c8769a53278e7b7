#!/bin/bash
set -x
mkdir -p gpurun_out
for tool in memcheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_probe.py infer 2>&1 | grep -v "^$" | tail -14 | cut -c1-220 | sed "s/^/$tool /"
done
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_probe.py train 2>&1 | tail -6 | cut -c1-220
