# Development aid (run through gpurun on one B200): the GPU test-suite, smoke(), the default bench line.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_job_validate.sh'
python -m pytest tests -m gpu -x -q > gpurun_out/validate_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/validate_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 5 > gpurun_out/validate_bench_n1.json 2> gpurun_out/validate_bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d = json.load(open("gpurun_out/validate_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
for k, v in d["workloads"].items():
    print(" ", k, {a: b for a, b in v.items() if a in ("value", "unit", "ms", "error")})
PY
