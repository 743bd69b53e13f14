python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu13.log; cat gpurun_out/pytest_gpu13.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
