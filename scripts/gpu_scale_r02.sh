#!/bin/bash
# Round-2 scaling evidence on one 8-GPU box: bench.py at N = 2, 4, 8 for the three workloads (N = 1 lines come from the
# single-GPU runs).  One JSON line per run, appended to gpurun_out/r02_scale.jsonl.
set -u
mkdir -p gpurun_out
out=gpurun_out/r02_scale.jsonl
: > $out
run() {  # n, args...
  n=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n "$@" 2> gpurun_out/scale_err_$n.log | grep '^{' >> $out
  echo "N=$n $* exit ${PIPESTATUS[0]}"
}
for n in 2 4 8; do run $n --steps 10 --warmup 3; done
for n in 2 8; do run $n --workload cfg5 --steps 2 --warmup 1; done
for n in 2 8; do run $n --workload cfg3_train --steps 10 --warmup 3; done
run 8 --workload cfg3_train --steps 10 --warmup 3 --graph
python - <<'PY'
import json
for l in open('gpurun_out/r02_scale.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'][:14], 'N', d['n_gpus'], 'value %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e', ('%.4g' % d['e2e']['value']) if d.get('e2e') else None)
PY
