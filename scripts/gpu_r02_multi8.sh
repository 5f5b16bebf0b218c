#!/bin/bash
# round 2, final 8-GPU lines (weak = the stated chain count per GPU, strong = in total)
N=8; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
brief() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  n_gpus %d value %.4g e2e %.4g ms/step %.1f frac %.3f bad %d procs %s coll %s' % (d['n_gpus'], d['value'], d['e2e']['value'] if 'e2e' in d else 0, d['ms_per_step'], d['roofline']['frac'], d['chains_with_error_status'], d['config'].get('processes'), (d.get('collective') or {}).get('backend')))
    elif 'rror' in l or 'Traceback' in l: print(l.rstrip()[-300:])
"; }
run() { label=$1; shift; echo "== $label"; timeout 600 "$@" 2> gpurun_out/r02_multi_err.txt | tee -a gpurun_out/r02_scale_n$N.jsonl | brief; grep -i -E "error|Traceback" gpurun_out/r02_multi_err.txt | tail -n 3; }
: > gpurun_out/r02_scale_n$N.jsonl
run "c3 weak torchrun x8" $TR bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline
run "c3 strong torchrun x8" $TR bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline --scaling strong
run "c3 weak one handle x8" python bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline
run "c1 weak torchrun x8" $TR bench.py --gpus $N --workload c1 --steps 3 --warmup 3 --no-cpu-baseline
run "c4 pooled weak torchrun x8" $TR bench.py --gpus $N --workload c4 --steps 3 --warmup 2 --no-cpu-baseline
run "c4 pooled weak one handle x8 (in-process NCCL)" python bench.py --gpus $N --workload c4 --steps 3 --warmup 2 --no-cpu-baseline
run "c5 pooled strong torchrun x8" $TR bench.py --gpus $N --workload c5 --steps 2 --warmup 2 --no-cpu-baseline --scaling strong
run "c2 strong torchrun x8" $TR bench.py --gpus $N --workload c2 --steps 4 --warmup 3 --no-cpu-baseline --scaling strong
