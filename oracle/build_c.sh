#!/bin/bash
# Build the oracle's C restatement (checker / CPU baseline only) -> oracle/_build/
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p _build
gcc -O3 -march=native -fopenmp -shared -fPIC sem_oracle_c.c -o _build/libsem_oracle_c.so
echo "built $(pwd)/_build/libsem_oracle_c.so"
