#!/bin/bash
# N-GPU ablation table of DESIGN.md section 6: usage tools/scale_ablation.sh N [out.jsonl]
N=${1:-8}; OUT=${2:-gpurun_out/scale_ablation_n$N.jsonl}; PORT=29511
run() { tag=$1; shift; echo "== $tag" >&2; line=$(env "${ENVV[@]}" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --quick --steps 8 "$@" 2>/dev/null | grep '"quick"' | tail -1); echo "{\"tag\": \"$tag\", \"line\": $line}" | tee -a $OUT; PORT=$((PORT+1)); }
: > $OUT
ENVV=(A=1); run default --with-e2e
ENVV=(A=1); run gather_off --gather off
ENVV=(A=1); run gather_padded_all_static --gather padded --submap all
ENVV=(A=1); run no_pin --no-pin --with-e2e
ENVV=(SCVOD_TRACK_PRIORITY=0); run flat_track_priority
nproc >&2
