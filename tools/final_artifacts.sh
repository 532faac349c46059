set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/gputest_r02e.txt; cat gpurun_out/gputest_r02e.txt
python bench.py > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err
tail -c 400 gpurun_out/bench_r02e.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02e_unet_step_launches.csv python tools/profile_step.py unet 2 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hop_chain --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02e_chain python tools/run_terms_once.py 64 32 64 4 > gpurun_out/ncu_chain.log 2>&1
ncu --set full --clock-control none -k regex:"mix_tma|wgrad_tma" --launch-skip 3 --launch-count 3 -f -o gpurun_out/r02e_dense python tools/profile_step.py layer 2 32 256 128 > gpurun_out/ncu_dense.log 2>&1
python tools/step_kineto.py 3 > gpurun_out/r02e_kineto_pdl.txt 2>/dev/null
DSW_OPTIONS=15=1 python tools/step_kineto.py 3 > gpurun_out/r02e_kineto_nopdl.txt 2>/dev/null
python tools/bench_layers.py > gpurun_out/r02e_layers.txt 2>/dev/null
ls -la gpurun_out | head -40
