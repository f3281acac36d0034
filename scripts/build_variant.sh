#!/bin/bash
# usage: build_variant.sh NAME "<extra nvcc flags>": builds the library with the flags into variants/NAME.so (A/B runs on the GPU box)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
LS_NVCC_EXTRA="$2" python -m livingscenes_b200._build > /dev/null
cp livingscenes_b200/_ls_b200.so variants/$1.so
echo "built variants/$1.so with [$2]"
