#!/bin/bash
# 2-GPU sanity of the multi-process paths after kernel changes: cfg2, cfg4 and the cfg3 training step (eager and graph).
set -u
mkdir -p gpurun_out
out=gpurun_out/scale2_check.jsonl
: > $out
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus 2 "$@" 2> gpurun_out/scale2_err.log | grep '^{' >> $out
  echo "N=2 $* exit ${PIPESTATUS[0]}"
}
run --steps 5 --warmup 3
run --workload cfg4 --steps 10 --warmup 3
run --workload cfg3_train --steps 10 --warmup 3
run --workload cfg3_train --steps 10 --warmup 3 --graph
python - <<'PY'
import json
for l in open('gpurun_out/scale2_check.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'][:14], 'N', d['n_gpus'], 'value %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e', ('%.4g' % d['e2e']['value']) if d.get('e2e') else None, (d.get('sample_direction') or {}).get('value'))
PY
