#!/bin/bash
# GPU session: the whole -m gpu suite, both bench arms (default line), the launch list of one timed step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f e2e %.4g solve %.4f frac %.3f"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['avg_launch_ms'],d['roofline']['frac']))
print({k:round(v['ms'],4) for k,v in d['kernels'].items()})
for k,v in (d['other_workloads'] or {}).items(): print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk not in ('config','kernels')})
PY
