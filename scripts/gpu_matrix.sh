#!/bin/bash
for ew in ${EWS:-16}; do for c in ${MODES:-1}; do for ck in ${CHUNKS:-32}; do for dbg in ${DBG:-0}; do
echo "EW=$ew"; FC_LINEAR_EW=$ew FC_LINEAR_MODE=$c FC_LINEAR_CHUNK_K=$ck FC_LINEAR_DEBUG=$dbg timeout 120 python scripts/bench_linear.py 2>&1 | tail -${TAILN:-2}
done; done; done; done
