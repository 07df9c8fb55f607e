#!/bin/bash
# Same-box A/B of compile-time variants: tools/ab_build.sh "<nvcc flags A>" "<nvcc flags B>" ...  (run on the GPU box)
# Rebuilds libmfsr_b200.so with each flag set and prints the bench's stage times.
for flags in "$@"; do
  MFSR_NVCC_EXTRA="$flags" python -m multi_frame_super_resolution_b200.build --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  for rep in 1 2; do
    MFSR_NVCC_EXTRA="$flags" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$flags]', d['value'], 'merge', d['roofline']['ms_per_launch'], 'flow', d['stage_ms']['flow'], 'total', d['ms_per_step'])"
  done
done
MFSR_NVCC_EXTRA="" python -m multi_frame_super_resolution_b200.build --force > /dev/null 2>&1
