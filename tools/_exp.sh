timeout 900 python bench.py > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; tail -c 300 gpurun_out/bench_v8.err
GKR_TRACE=1 timeout 200 python bench.py --steps 3 --warmup 3 --sumcheck-vars 0 --large-layer-k 0 --no-cpu 2>&1 >/dev/null | grep "gkr trace" | tail -4 | head -2 > gpurun_out/host_trace.txt
GKR_NO_PRELAUNCH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_v2.csv python bench.py --steps 1 --warmup 1 --layers 2 --no-cpu --sumcheck-vars 0 --large-layer-k 0 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-200; wc -l gpurun_out/launches_v2.csv
