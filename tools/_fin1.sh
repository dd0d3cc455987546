mkdir -p gpurun_out/fin3
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/fin3/bench_n1.json 2> gpurun_out/fin3/bench_n1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/fin3/bench_ref.json 2> gpurun_out/fin3/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/fin3/launches.csv python bench.py --steps 2 --warmup 3 --seeds 1 --sumcheck-vars 26 --sumcheck-vars-small 0 --large-layer-k 0 --tcircom-inputs 0 --no-cpu > gpurun_out/fin3/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/fin3/bench_n1.json'));print(d['value'],d['seeds'],d['host'],d['t_circom_like_batch']['proofs_per_s'])"
