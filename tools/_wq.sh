mkdir -p gpurun_out/wir
timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -4 > gpurun_out/wir/pytest3.txt
cat gpurun_out/wir/pytest3.txt
F="--steps 10 --warmup 3 --no-cpu --sumcheck-vars 0 --sumcheck-vars-small 0 --large-layer-k 0 --seeds 1 --tcircom-inputs 0"
timeout 300 python bench.py $F > gpurun_out/wir/b3.json 2> gpurun_out/wir/b3.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/wir/b3.json"))
print(round(d["value"],3), d["gpu_launches"], d.get("parity",{}).get("all_ok"))
PY
GRID=1x16,4x16,16x4 timeout 200 python tools/batch_scaling.py 64 native | cut -c1-200
which perf gdb valgrind 2>&1 | head -3
