python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -x -q 2>&1 | tail -1
for sh in "32 32 24 4" "64 32 64 4" "32 32 32 4"; do python tools/time_terms.py $sh 2>/dev/null; DSW_LIB_PATH=/root/repo/deepsphere-weather_b200/libdsw_prev.so python tools/time_terms.py $sh 2>/dev/null; done
python tools/time_step.py 20 2>/dev/null; DSW_LIB_PATH=/root/repo/deepsphere-weather_b200/libdsw_prev.so python tools/time_step.py 20 2>/dev/null
