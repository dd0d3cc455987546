for g in 0 1 0 1; do
if [ $g = 1 ]; then export GKR_AUX_GATE=1; else unset GKR_AUX_GATE; fi
timeout 200 python bench.py --steps 8 --warmup 3 --sumcheck-vars 0 --large-layer-k 0 --no-cpu > gpurun_out/bench_tr.json 2> gpurun_out/bench_tr.err
python -c "import json,sys; d=json.load(open('gpurun_out/bench_tr.json')); print('gate=$g', d['value'], d['e2e']['value'], d['host'])"
done
