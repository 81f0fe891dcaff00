# development aid: full GPU test-suite + smoke (one B200)
python -m pytest tests -m gpu -x -q > gpurun_out/s22_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/s22_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
