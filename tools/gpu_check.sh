timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/latest_bench_nocpu.json 2> gpurun_out/latest_bench.err
tail -3 gpurun_out/latest_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/latest_bench_nocpu.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["stream_ordered"], d["large_batch"]["value"], d["e2e"]["value"])
PY
