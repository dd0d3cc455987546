timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for ns in 0 1 0 1; do
if [ $ns = 1 ]; then export GKR_NO_EQ_SPLIT=1; else unset GKR_NO_EQ_SPLIT; fi
timeout 300 python bench.py --steps 8 --warmup 3 --sumcheck-vars 0 --large-layer-k 0 --tcircom-inputs 0 --no-cpu > gpurun_out/bench_tr.json 2> gpurun_out/bench_tr.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_tr.json').read().strip().splitlines()[-1]); print('nosplit=$ns', d['value'], d['e2e']['value'], d['host'], {k:round(v['ms'],2) for k,v in d.get('kernel_classes',{}).items() if k in ('wiring','eq')})"
done
