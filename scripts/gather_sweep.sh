#!/usr/bin/env bash
# Sweep of scripts/gather_stream_bench.cu on the GPU box (binary prebuilt: scripts/gather_stream_bench.bin).
B=scripts/gather_stream_bench.bin
OUT=gpurun_out/r2f_gather_sweep.log
: > $OUT
run() { echo "== $*" >> $OUT; timeout 60 $B "$@" 2>&1 | grep -v watchdog >> $OUT || echo "FAILED rc=$?" >> $OUT; }
# mode stages nprod shift bload transform rot ctas noload nostore nissue promo storemode bmode Wp ld
run 2 4 4 384 0 1 0 1 0 0 4 256 1 0 384
run 2 4 4 384 0 1 0 1 0 0 4 256 0 0 384
run 2 4 4 384 1 1 0 1 0 0 4 256 1 0 384
run 2 4 4 384 1 1 0 1 0 0 4 256 0 0 384
run 2 4 4 384 1 1 0 1 0 0 4 256 1 1 384
run 2 3 4 384 1 1 0 1 0 0 4 256 1 0 384
run 2 4 4 0 1 1 0 1 0 0 4 256 1 0 384
run 2 4 4 0 1 1 0 1 0 0 4 256 0 0 384
run 2 4 4 384 0 0 0 1 1 0 4 256 1 0 384
run 2 4 4 384 0 0 0 1 1 0 4 256 0 0 384
run 2 4 4 384 0 0 0 1 0 1 4 256 1 0 384
run 2 4 4 372 1 1 0 1 0 0 4 256 1 0 372 30880
run 2 4 4 0 1 1 0 1 0 0 4 256 1 0 372 30880
cat $OUT
