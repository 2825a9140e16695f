#!/usr/bin/env bash
# compute-sanitizer over one small pass of every operator (the smoke test) -- memcheck and racecheck.
# Run on the GPU box:  gpurun -- 'bash tools/sanitize.sh'   -> gpurun_out/sanitizer_{memcheck,racecheck}.txt
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool exit code: $?" >> gpurun_out/sanitizer_$tool.txt
  tail -4 gpurun_out/sanitizer_$tool.txt
done
