#!/bin/sh
# Host-emulation build of the kernels (test infrastructure; see csrc/rhccq_common.cuh).
set -e
here=$(cd "$(dirname "$0")" && pwd)
src=$here/../../roibasedimagecompression_b200/csrc
out=$here/_build
mkdir -p "$out"
g++ -O2 -g -std=c++17 -fPIC -shared -ffp-contract=off -DRHCCQ_HOST_EMU -x c++ \
    "$src/rhccq_api.cu" "$src/rhccq_palette.cu" "$src/rhccq_split.cu" "$src/rhccq_minibatch.cu" "$src/rhccq_points.cu" "$src/rhccq_pixels.cu" "$src/rhccq_merge.cu" \
    -o "$out/librhccq_emu.so"
echo "$out/librhccq_emu.so"
