#!/bin/sh
# Host-emulation build of the kernels (test infrastructure; see csrc/rhccq_common.cuh).
set -e
here=$(cd "$(dirname "$0")" && pwd)
src=$here/../../roibasedimagecompression_b200/csrc
out=$here/_build
mkdir -p "$out"
g++ -O2 -g -std=c++17 -fPIC -shared -ffp-contract=off -mfma -DRHCCQ_HOST_EMU -x c++ \
    "$src/rhccq_api.cu" "$src/rhccq_palette.cu" "$src/rhccq_split.cu" "$src/rhccq_minibatch.cu" "$src/rhccq_points.cu" "$src/rhccq_pixels.cu" "$src/rhccq_merge.cu" "$src/rhccq_deflate.cu" \
    -o "$out/librhccq_emu.so" &
# the same with every K-Means decision forced through its float64 re-evaluation (csrc/rhccq_split.cu)
g++ -O2 -g -std=c++17 -fPIC -shared -ffp-contract=off -mfma -DRHCCQ_HOST_EMU -DRHCCQ_KM_FORCE_EXACT -x c++ \
    "$src/rhccq_api.cu" "$src/rhccq_palette.cu" "$src/rhccq_split.cu" "$src/rhccq_minibatch.cu" "$src/rhccq_points.cu" "$src/rhccq_pixels.cu" "$src/rhccq_merge.cu" "$src/rhccq_deflate.cu" \
    -o "$out/librhccq_emu_exact.so" &
wait
echo "$out/librhccq_emu.so"
