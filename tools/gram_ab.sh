#!/bin/bash
# tools/gram_ab.sh build | run  -- A/B of Gram-kernel variants (compile-time knobs of csrc/gram.cu) on one B200.
#   build (here, no GPU): one libobboot.so per variant under tools/_ab/<name>/ (only gram.o differs; the other objects
#                         are the ones `make` left in oaxaca_blinder_rs_b200/_lib/obj) + the probe binary
#   run   (on the GPU box): the probe against every variant, results to gpurun_out/gram_ab.log
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
AB="$ROOT/tools/_ab"
CSRC="$ROOT/oaxaca_blinder_rs_b200/csrc"
OBJ="$ROOT/oaxaca_blinder_rs_b200/_lib/obj"
# VARIANTS="name:-DFLAG=..,-DFLAG2=.. name2:..." -- compile-time knobs of an experimental gram.cu (the round-2 hand-off
# experiment, profiles/r02_gram_handoff_ab.log, used OB_GRAM_XSTAGE / OB_GRAM_DEPHASE; those variants are gone again)
VARIANTS=${VARIANTS:-"base:-DOB_AB_BASE=1"}
case "$1" in
build)
    mkdir -p "$AB"
    for v in $VARIANTS; do
        name=${v%%:*}; flags=$(echo "${v#*:}" | tr ',' ' ')
        mkdir -p "$AB/$name"
        ( cd "$CSRC" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $flags \
              -c gram.cu -o "$AB/$name/gram.o" 2> "$AB/$name/gram.ptxas.log" \
          && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$AB/$name/libobboot.so" "$AB/$name/gram.o" \
              $(ls "$OBJ"/*.o | grep -v '/gram.o$') -cudart static -ldl -lpthread ) &
    done
    wait
    gcc -O2 -I "$ROOT/include" "$ROOT/tools/gram_ab.c" -L "$ROOT/oaxaca_blinder_rs_b200/_lib" -lobboot -lm -o "$AB/gram_ab"
    ls -la "$AB"/*/libobboot.so "$AB/gram_ab"
    ;;
run)
    shift
    mkdir -p "$ROOT/gpurun_out"
    LOG="$ROOT/gpurun_out/${LOGNAME_AB:-gram_ab.log}"
    : > "$LOG"
    for v in $VARIANTS; do
        name=${v%%:*}
        echo "== $name (${v#*:})" >> "$LOG"
        LD_LIBRARY_PATH="$AB/$name" timeout 120 "$AB/gram_ab" "$@" >> "$LOG" 2>&1 || echo "FAILED rc=$?" >> "$LOG"
    done
    cat "$LOG"
    ;;
*) echo "usage: $0 build | run [n p reps runs]"; exit 2 ;;
esac
