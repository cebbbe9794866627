#!/bin/bash
# experiment: pick_sparse at different register / occupancy targets (run on the GPU box)
for mb in 4 5 6; do
  touch gateway-api-inference-extension_b200/csrc/pick_sparse.cu
  EPPSCORE_NVCC_EXTRA="-DEPP_SPARSE_MINBLOCKS=$mb" python gateway-api-inference-extension_b200/build.py > /dev/null
  echo "== EPP_SPARSE_MINBLOCKS=$mb"
  EPPSCORE_NVCC_EXTRA="-DEPP_SPARSE_MINBLOCKS=$mb" timeout 200 python tools/prof_step.py pick 2>&1 | grep "pick_sparse"
done
touch gateway-api-inference-extension_b200/csrc/pick_sparse.cu
python gateway-api-inference-extension_b200/build.py > /dev/null
