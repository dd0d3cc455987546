F="--steps 16 --warmup 4 --no-cpu --sumcheck-vars 0 --sumcheck-vars-small 0 --large-layer-k 0 --tcircom-inputs 0 --seeds 1"
timeout 200 python bench.py $F > /dev/null 2>&1
for g in 11 9 8 7 6 11 9 8 7 6; do
GKR_LINE_TAIL_LOG2=$g GKR_TRACE=1 timeout 200 python bench.py $F > /tmp/b.json 2> /tmp/b.err
echo "tail $g: $(python -c "import json;d=json.load(open('/tmp/b.json'));print(round(d['value'],3), d['parity'].get('all_ok'))") $(grep 'gkr trace\] wait_direct=[1-9]' /tmp/b.err | tail -1 | grep -o 'aux_sync=[0-9.]*ms')"
done
timeout 300 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_prove.py -q -x -m gpu 2>&1 | tail -2
