#!/bin/bash
# usage: tools/variant_probe2.sh lib1.so lib2.so ...  (the default build first) -> Granger stage timings per variant
cd "$(dirname "$0")/.."
for lib in "" "$@"; do
  echo "== ${lib:-default build}"
  SC_B200_LIB=${lib:+$(realpath $lib)} python tools/granger_probe2.py 2>&1 | tail -4
done
