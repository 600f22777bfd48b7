#!/bin/bash
# usage: profiles/run_multi.sh N   -- the multi-GPU measurements of one box with N GPUs (all under timeout):
#   bit-identity of both transports against one engine, the default bench line (peer-memory transport; it
#   carries BASELINE configs[3] / [4] and, at N = 8, the north_star configuration as sub-records), and the
#   same weak-scaling point on the NCCL transport
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
    print('  value %.4g /s  %.4f ms/step  e2e %s  parity %s  clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('parity_check') or {}).get('bit_identical'), d['clocks']))
    print('  phases', d.get('slab_phases_ms_per_step'))
    if 'north_star_16m' in d: print('  north_star_16m %.4g /s  %.4f ms/step' % (d['north_star_16m']['value'], d['north_star_16m']['ms_per_step']))
    for k, v in d.get('other_configs', {}).items(): print('  %s %.4g /s  %.4f ms/step' % (k, v['value'], v['ms_per_step']))
"; grep -v "OMP_NUM\|\*\*\*\|^$\|Warning\|warn" $OUT/$name.err | tail -3
}
timeout 300 $TR tests/run_slab_nccl.py 2>&1 | grep "bit_identical"
run r02_bench_${N}gpu_p2p 500 --steps 50
run r02_bench_${N}gpu_nccl 300 --steps 50 --transport nccl --no-north-star --no-parity-check
