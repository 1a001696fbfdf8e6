#!/bin/bash
# The CPU oracle (test infrastructure) under AddressSanitizer + UBSan: rebuilds oracle/liboracle.so with the sanitizers, runs the
# oracle's CPU tests with the runtimes preloaded into python, and restores the normal build whatever happens.
#   bash tools/fuzz/sanitize_oracle.sh
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
GXX="${SAN_CXX:-/usr/bin/g++}"          # a g++ that ships libasan / libubsan
restore() { make -C "$ROOT/oracle" -B all >/dev/null 2>&1; }
trap restore EXIT
make -C "$ROOT/oracle" -B all CXX="$GXX" CXXFLAGS="-std=c++17 -O1 -g -ffp-contract=off -fPIC -pthread -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer" >/dev/null || exit 1
cd "$ROOT"
LD_PRELOAD="$("$GXX" -print-file-name=libasan.so) $("$GXX" -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0 \
UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
python -m pytest tests/test_oracle_builder.py tests/test_oracle_traversal.py tests/test_raygen_oracle.py tests/test_get_data.py tests/test_collide.py tests/test_materials.py \
    -x -q -m "not gpu" -p no:cacheprovider
