#!/bin/bash
# ASan + UBSan mutation fuzzing of the host-side OBJ / glTF / GLB readers (no GPU needed):
#   bash tools/fuzz/run_fuzz_loader.sh [iterations] [rng seed] [work dir]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
N="${1:-20000}"; SEED="${2:-1}"; DIR="${3:-/tmp/cndl_fuzz}"
mkdir -p "$DIR"
python "$ROOT/tools/fuzz/make_seeds.py" "$DIR" >/dev/null
g++ -std=c++17 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -x c++ "$ROOT/candela_b200/csrc/model_loader.cu" "$ROOT/tools/fuzz/fuzz_loader.cpp" -o "$DIR/fuzz_loader"
ASAN_OPTIONS=detect_leaks=1:allocator_may_return_null=1:max_allocation_size_mb=2048 UBSAN_OPTIONS=print_stacktrace=1 "$DIR/fuzz_loader" "$DIR" "$N" "$SEED"
