#!/bin/bash
# usage: profiles/run_multi.sh N   -- the multi-GPU measurements of one box with N GPUs (all under timeout)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
OUT=gpurun_out
run() { # name, timeout, args...
  name=$1; to=$2; shift 2
  timeout $to $TR bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err
  echo "== $name rc=$?"; grep '^{' $OUT/$name.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('  value %.4g /s  %.4f ms/step  e2e %s  parity %s  transport %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('parity_check') or {}).get('bit_identical'), d['config'].get('note','')[-60:]))
    if 'north_star_16m' in d: print('  north_star_16m %.4g /s  %.4f ms/step' % (d['north_star_16m']['value'], d['north_star_16m']['ms_per_step']))
"; grep -v "OMP_NUM\|\*\*\*\|^$\|Warning\|warn" $OUT/$name.err | tail -3
}
timeout 300 $TR tests/run_slab_nccl.py 2>&1 | grep "bit_identical"
run r02_bench_${N}gpu_p2p 300 --steps 50
if [ "$N" = "8" ]; then
  run r02_bench_8gpu_mixed16m 300 --workload mixed3d --particles 2000000 --steps 20 --no-north-star
fi
run r02_bench_${N}gpu_128m 400 --particles 16000000 --steps 10 --warmup 3 --no-north-star --no-parity-check --e2e-steps 0
run r02_bench_${N}gpu_nccl 300 --steps 50 --transport nccl --no-north-star --no-parity-check
