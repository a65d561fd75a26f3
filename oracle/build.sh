#!/bin/sh
# Builds the CPU oracle (test infrastructure, never shipped in the product path).
# Output: oracle/_build/liboracle.so
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/_build"
gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC -std=c99 \
    -o "$here/_build/liboracle.so" "$here/oracle.c" -lm
