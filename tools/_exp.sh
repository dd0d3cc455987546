for wk in 6 8 12 24; do
GKR_BATCH_WORKERS=$wk timeout 900 python bench.py --steps 1 --warmup 3 --sumcheck-vars 0 --large-layer-k 0 --no-cpu > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_tc.json').read().strip().splitlines()[-1]); t=d['t_circom_like_batch']; print($wk, t['ms_total'], t['proofs_per_s'], t['ms_one_input_alone'])"
done
