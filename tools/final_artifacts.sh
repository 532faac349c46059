set -x
python bench.py > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err
tail -c 400 gpurun_out/bench_r02d.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02d_unet_step_launches.csv python tools/profile_step.py unet 2 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hop_chain --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02d_chain python tools/run_terms_once.py 64 32 64 4 > gpurun_out/ncu_chain.log 2>&1
ncu --set full --clock-control none -k regex:"mix_tma|wgrad_tma" --launch-skip 3 --launch-count 3 -f -o gpurun_out/r02d_dense python tools/profile_step.py layer 2 32 256 128 > gpurun_out/ncu_dense.log 2>&1
python tools/step_kineto.py 3 > gpurun_out/r02d_kineto_pdl.txt 2>/dev/null
DSW_OPTIONS=15=1 python tools/step_kineto.py 3 > gpurun_out/r02d_kineto_nopdl.txt 2>/dev/null
python tools/bench_layers.py > gpurun_out/r02d_layers.txt 2>/dev/null
ls -la gpurun_out | head -40
