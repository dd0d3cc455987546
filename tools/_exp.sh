timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for lg in 18 19 20; do
GKR_LOOKAHEAD_LOG2=$lg timeout 200 python bench.py --steps 8 --warmup 3 --sumcheck-vars 0 --large-layer-k 0 --no-cpu > gpurun_out/bench_tr.json 2> gpurun_out/bench_tr.err
python -c "import json,sys; d=json.load(open('gpurun_out/bench_tr.json')); print($lg, d['value'], d['e2e']['value'], d['host'])"
done
GKR_TRACE=1 GKR_LOOKAHEAD_LOG2=20 timeout 200 python bench.py --steps 4 --warmup 3 --sumcheck-vars 0 --large-layer-k 0 --no-cpu > gpurun_out/bench_tr.json 2> gpurun_out/bench_tr.err
grep "gkr trace" gpurun_out/bench_tr.err | tail -4 | head -2 | cut -c1-330
