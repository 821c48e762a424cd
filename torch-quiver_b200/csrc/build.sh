#!/usr/bin/env bash
# Build libquiver_b200.so (sm_100a only) next to the Python adapter.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../torch_quiver/libquiver_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$HERE/build"
objs=()
pids=()
for f in qv_runtime qv_xorwow qv_sample qv_gather; do
  src="$HERE/$f.cu"; obj="$HERE/build/$f.o"
  stale=0
  for dep in "$src" "$HERE"/*.cuh "$ROOT/include/quiver_b200.h"; do
    if [[ ! -f "$obj" || "$dep" -nt "$obj" ]]; then stale=1; fi
  done
  if [[ $stale == 1 ]]; then
    "$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden \
      -I"$ROOT/include" "$@" -c "$src" -o "$obj" &
    pids+=($!)
  fi
  objs+=("$obj")
done
for pid in "${pids[@]:-}"; do
  if [[ -n "$pid" ]]; then wait "$pid" || { echo "build failed" >&2; rm -f "$HERE"/build/*.o; exit 1; }; fi
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC "${objs[@]}" -o "$OUT" -lpthread -ldl -lrt
echo "built $OUT"
