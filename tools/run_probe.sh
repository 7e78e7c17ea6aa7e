#!/bin/bash
# runs the operand-convention probe matrix on a B200 (each run bounded: a bad descriptor must not hang the box)
P=tools/bin/umma_probe
run() { timeout 20 $P "$@" || echo "FAILED/timeout: $*"; }
for img in 0 1 2; do run 0 3 64 128 $img; done
run 0 3 6 128 1
for img in 0 1 2 3; do run 1 3 64 128 $img; run 1 3 16 128 $img; done
for img in 0 1 2; do run 2 1 64 128 $img; done
for img in 0 1 2; do run 3 1 64 128 $img; run 3 3 64 128 $img; done
run 3 3 64 64 1
run 3 3 64 32 1
