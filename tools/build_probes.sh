#!/bin/bash
# Builds the stand-alone GPU probes under tools/ into tools/bin/ (git-ignored; ships to the GPU box with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
for src in tools/probe_*.cu; do
  name=$(basename "$src" .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I rule_guided_music_b200/csrc "$src" -o "tools/bin/$name" -lcuda
done
ls -la tools/bin
