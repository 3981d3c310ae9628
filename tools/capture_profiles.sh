#!/bin/bash
# One GPU call that produces the ncu evidence kept under profiles/ (round tag = $1, default r02). Every step has its own timeout.
TAG=${1:-r02}; OUT=gpurun_out; mkdir -p $OUT
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python tools/run_batch.py 64 2 > $OUT/${TAG}_launches.log 2>&1
timeout 420 ncu --set full --import-source on --clock-control none -k regex:"^(?!k_track)" -c 40 -f -o $OUT/${TAG}_scan python tools/run_batch.py 64 1 > $OUT/${TAG}_scan.log 2>&1
timeout 240 ncu --set full --import-source on --clock-control none -k regex:"k_track" --launch-skip 30 -c 8 -f -o $OUT/${TAG}_track python tools/run_batch.py 64 1 > $OUT/${TAG}_track.log 2>&1
SCVOD_CHAIN_TMA=1 timeout 200 ncu --set full --import-source on --clock-control none -k regex:"k_patch_chain" -c 1 -f -o $OUT/${TAG}_chain_tma python tools/run_batch.py 64 1 > $OUT/${TAG}_chain_tma.log 2>&1
ls -la $OUT
