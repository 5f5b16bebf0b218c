#!/bin/bash
timeout 900 python -m pytest tests/test_full_size.py -m gpu -q > gpurun_out/r02_gputest_fullsize.log 2>&1; tail -n 15 gpurun_out/r02_gputest_fullsize.log
timeout 600 python bench.py --steps 5 --warmup 3 --dump-stride 10 --no-cpu-baseline > gpurun_out/r02_bench_c3_dump10.json 2> gpurun_out/r02_bench_c3_dump10.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_dump10.json')); print('dump10 value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d['dumps']['snapshots_popped'], d['dumps']['snapshots_enqueued'], d['gpu_launches'])"
