#!/bin/sh
# ThreadSanitizer check of the samplers' verifier thread (host code only: sampler.cu is compiled as
# C++ against a shim of common.cuh).  Dense lists + MACR_SAMPLER_THREADS=2: a rewind every few
# triples.  Prints "equal triples: 1  equal state: 1" and no ThreadSanitizer report when clean.
set -e
here=$(cd "$(dirname "$0")" && pwd)
tmp=$(mktemp -d)
cp "$here/common.cuh" "$here/main.cpp" "$tmp/"
cp "$here/../../macr_b200/csrc/sampler.cu" "$tmp/sampler.cpp"
g++ -O1 -g -std=c++17 -fsanitize=thread -msse2 -w -o "$tmp/tsan_test" "$tmp/sampler.cpp" "$tmp/main.cpp" -lpthread
"$tmp/tsan_test"
rm -rf "$tmp"
