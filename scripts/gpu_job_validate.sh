# development aid: targeted GPU tests + bench (one B200)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "gram or qr or tensordot or tall or tsqr or config3" > gpurun_out/s11_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/s11_pytest.log
python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/s11_bench_n1.json 2> gpurun_out/s11_bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/s11_bench_n1.json"))
print(d["value"], d["ms_per_step"])
for k,v in d["workloads"].items(): print(" ",k,{a:b for a,b in v.items() if a in ("value","unit","ms","error")})
PY
