# development aid: glms drop-in with multi-block launches (one B200)
python -m pytest tests/test_gpu_reference_api.py tests/test_gpu_parity.py -x -q -k "glms or lr or newton or fused" > gpurun_out/s16_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/s16_pytest.log
python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/s16_bench_n1.json 2> gpurun_out/s16_bench_n1.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open("gpurun_out/s16_bench_n1.json"))
for k,v in d["workloads"].items(): print(" ",k,{a:b for a,b in v.items() if a in ("value","unit","ms","error")})
PY
