mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log 2>&1
python bench.py > gpurun_out/bench_n1.log 2>&1
python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
cat gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; tail -1 gpurun_out/bench_n1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic']); print(json.dumps(d['other_workloads'], indent=0)[:1800])"
