#!/bin/bash
# TEST INFRASTRUCTURE: compile the unmodified CUDA sources of channelflow_b200/csrc for the CPU emulator
# (tests/emu/cuda_runtime.h) into tests/_emu/libcfgpu_emu.so.  Used only by `pytest -m "not gpu"`.
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
OUT="$ROOT/tests/_emu"
mkdir -p "$OUT/obj"
CXX=${CXX:-g++}
FLAGS="-O2 -g -std=c++17 -fPIC -DCF_EMU -I$ROOT/tests/emu -I$ROOT/channelflow_b200/csrc -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable -Wno-sign-compare"
pids=()
for f in "$ROOT"/channelflow_b200/csrc/*.cu; do
  o="$OUT/obj/$(basename "$f" .cu).o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find "$ROOT/channelflow_b200/csrc" "$ROOT/tests/emu" "$ROOT/include" -newer "$o" \( -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
    $CXX $FLAGS -x c++ -c "$f" -o "$o" &
    pids+=($!)
  fi
done
o="$OUT/obj/cfemu.o"
if [ ! -f "$o" ] || [ "$ROOT/tests/emu/cfemu.cpp" -nt "$o" ] || [ "$ROOT/tests/emu/cuda_runtime.h" -nt "$o" ]; then
  $CXX $FLAGS -c "$ROOT/tests/emu/cfemu.cpp" -o "$o" &
  pids+=($!)
fi
for p in "${pids[@]}"; do wait $p; done
if [ ! -f "$OUT/libcfgpu_emu.so" ] || [ -n "$(find "$OUT/obj" -name '*.o' -newer "$OUT/libcfgpu_emu.so" | head -1)" ]; then
  $CXX -shared -Wl,-Bsymbolic -o "$OUT/libcfgpu_emu.so" "$OUT"/obj/*.o -lpthread
fi
echo "built $OUT/libcfgpu_emu.so"
