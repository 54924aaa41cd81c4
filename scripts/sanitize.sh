#!/bin/bash
# Race / memory-error check of the C++ core (protocol, sync engine, fabric, RPC, loaders) on the CPU backend:
# builds the native application with ThreadSanitizer and with AddressSanitizer + UBSan (CUDA entry points stubbed)
# and runs the reference-scale dynamic-allocation stress (threads as ranks: relocation + replication under fully
# asynchronous pushes). Exit code 0 = no report.   bash scripts/sanitize.sh [build-dir]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/adapm_b200/csrc
OUT=${1:-/tmp/adapm_sanitize}
mkdir -p "$OUT"
FILES="$SRC/apps/cuda_stub.cc $SRC/apps/simple.cc $SRC/adapm/fabric.cc $SRC/adapm/node.cc $SRC/adapm/rpc.cc \
       $SRC/adapm/sampling.cc $SRC/adapm/store_cpu.cc $SRC/adapm/sync_engine.cc $SRC/adapm/corpus.cc $SRC/adapm/io.cc"
rc=0
for kind in thread address,undefined; do
  exe=$OUT/simple_${kind%%,*}
  g++ -std=c++17 -O1 -g -fsanitize=$kind -fno-omit-frame-pointer -Wno-tsan -I"$SRC" -o "$exe" $FILES -lpthread -lrt
  for args in "--stress 3000 -s 3 -t 2" "--fuzz 3000 -s 3 -t 2 --seed 3" "--fuzz 3000 -s 2 -t 2 --seed 4 --techniques relocation_only" "-k 10 -t 2 -i 3 -v 2"; do
    log=$OUT/$(basename "$exe")_$(echo "$args" | tr -c 'a-z0-9' _).log
    if ! TSAN_OPTIONS="halt_on_error=0" "$exe" $args > "$log" 2>&1; then rc=1; fi
    n=$(grep -c "WARNING: ThreadSanitizer\|ERROR: AddressSanitizer\|runtime error" "$log" || true)
    echo "sanitize=$kind args='$args': $n report(s)  ($log)"
    [ "$n" = "0" ] || rc=1
  done
done
exit $rc
