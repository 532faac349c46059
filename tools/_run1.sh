python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err
tail -c 300 gpurun_out/bench_r02d.json
