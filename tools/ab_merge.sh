#!/bin/bash
# Same-box A/B of compile-time variants of the merge kernel: tools/ab_merge.sh "<nvcc flags A>" "<nvcc flags B>" ...  (on the GPU box)
for flags in "$@"; do
  MFSR_NVCC_EXTRA="$flags" python -m multi_frame_super_resolution_b200.build --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "[$flags] $(MFSR_NVCC_EXTRA="$flags" python tools/merge_only.py 4 2>&1 | tail -1)"
done
MFSR_NVCC_EXTRA="" python -m multi_frame_super_resolution_b200.build --force > /dev/null 2>&1
