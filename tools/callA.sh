mkdir -p gpurun_out
(time timeout 420 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/A_pytest.log 2>&1
(time timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > gpurun_out/A_smoke.log 2>&1
(time timeout 150 python bench.py) > gpurun_out/A_bench.json 2> gpurun_out/A_bench.err
for dr in 1 0; do echo DEVICE_RANGE=$dr; ABL_CUDA_DEVICE_RANGE=$dr timeout 60 python tools/quick_slabs.py boids2d-1M-f64 --slabs 2 --transport direct --scale 2 --steps 200; done > gpurun_out/A_slabs.txt 2>&1
tail -5 gpurun_out/A_pytest.log; tail -3 gpurun_out/A_smoke.log; cat gpurun_out/A_bench.json | cut -c1-400; cat gpurun_out/A_slabs.txt
