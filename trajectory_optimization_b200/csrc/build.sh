#!/usr/bin/env bash
# Build libcovb200.so for sm_100a in-tree (the .so is git-ignored but travels with gpurun snapshots).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libcovb200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --shared
       -cudart static ${COV_PTXAS_V:+-Xptxas -v})
"${NVCC}" "${FLAGS[@]}" -o "${OUT}" "${HERE}/cov_api.cu" "${HERE}/cov_pose.cu" "${HERE}/cov_traj.cu" \
    "${HERE}/cov_tools.cu" "${HERE}/cov_sweep.cu" "${HERE}/cov_hull.cu" "${HERE}/cov_sort.cu" "${HERE}/cov_rig.cu" "${HERE}/cov_reg.cu" "${HERE}/cov_peer.cu"
echo "built ${OUT}"
