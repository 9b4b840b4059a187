#!/bin/bash
# tools/gram_tiles_ab.sh prepare <git-ref> | run  -- old vs new libobboot.so over problem shapes that exercise every tile
# class of the Gram kernel (full tiles, tail tiles of one, two and three quanta, partly filled tail panel, structural
# zeros).  The probe (tools/gram_ab.c) prints a hash of all standard errors, point statistics and CI bounds: equal
# hashes = bit-identical results.  This is how the quantum column tiling was checked (profiles/r02_gram_tiles_ab.log).
#   prepare <ref>  (here, no GPU): builds <ref> in a git worktree -> tools/_ab/base/, copies the current build ->
#                  tools/_ab/new/, compiles the probe
#   run            (on the GPU box): every case against both libraries -> gpurun_out/gram_tiles_ab.log
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
AB="$ROOT/tools/_ab"
case "$1" in
prepare)
    ref=${2:-HEAD}
    wt=$(mktemp -d /tmp/obboot_base.XXXXXX)
    git -C "$ROOT" worktree add -f "$wt" "$ref" -q
    make -C "$wt/oaxaca_blinder_rs_b200/csrc" -j8 -s
    make -C "$ROOT/oaxaca_blinder_rs_b200/csrc" -j8 -s
    mkdir -p "$AB/base" "$AB/new"
    cp "$wt/oaxaca_blinder_rs_b200/_lib/libobboot.so" "$AB/base/"
    cp "$ROOT/oaxaca_blinder_rs_b200/_lib/libobboot.so" "$AB/new/"
    git -C "$ROOT" worktree remove --force "$wt"
    gcc -O2 -I "$ROOT/include" "$ROOT/tools/gram_ab.c" -L "$ROOT/oaxaca_blinder_rs_b200/_lib" -lobboot -lm -o "$AB/gram_ab"
    ;;
run)
    cd "$ROOT"
    LOG=gpurun_out/gram_tiles_ab.log; mkdir -p gpurun_out; : > $LOG
    run() { for v in base new; do echo "== $v: $*" >> $LOG; LD_LIBRARY_PATH=$AB/$v timeout 100 $AB/gram_ab "$@" >> $LOG 2>&1 || echo "FAILED rc=$?" >> $LOG; done; }
    #   n        p  reps runs ncat      (K = 1 + p + 3 ncat)
    run 10000000 44 2000 2 2     # K = 51 with two 4-level categoricals: 11 tiles -> 10.75
    run 2000000 50 1000 2 0      # K = 51, no categoricals: 11 tiles both
    run 2000000 30 1000 2 0      # K = 31: 4.5 -> 4.25 tiles
    run 300000 16 300 1 0        # K = 17: 1.5 tiles both (two quanta)
    run 300000 11 300 1 0        # K = 12: 1 tile -> 3 quanta
    run 300000 5 300 1 0         # K = 6: half tile -> 1 quantum
    run 300000 6 300 1 1         # K = 10 with a categorical
    run 1000 3 77 1 1            # tiny, partly filled panel
    cat $LOG
    ;;
*) echo "usage: $0 prepare [git-ref] | run"; exit 2 ;;
esac
