set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_s3.log
tail -4 gpurun_out/pytest_s3.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_s3.log 2>&1; tail -2 gpurun_out/smoke_s3.log
bash tools/final_measure.sh s3
