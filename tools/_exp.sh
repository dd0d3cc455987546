timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
GKR_TRACE=1 timeout 200 python bench.py --steps 4 --warmup 3 --sumcheck-vars 0 --large-layer-k 0 --no-cpu > gpurun_out/bench_tr.json 2> gpurun_out/bench_tr.err
grep "gkr trace" gpurun_out/bench_tr.err | tail -4 | head -2 | cut -c1-330
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_tr.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['host'], {k:round(v['ms'],2) for k,v in d.get('kernel_classes',{}).items()})"
