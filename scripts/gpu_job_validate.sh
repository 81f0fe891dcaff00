python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_api.py -x -q -k "qr or tsqr or gram" > gpurun_out/s20_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/s20_pytest.log
python scripts/probe_qr_shifted.py 2>&1 | tail -4
NUMS_QR_SHIFTED=0 python scripts/probe_qr_shifted.py 2>&1 | tail -3
