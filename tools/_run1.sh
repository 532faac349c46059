python -m pytest tests/test_gpu_parity.py -x -q -k "hop or cheb or spmm or terms" 2>&1 | tail -1
for sh in "16 32 256 4" "8 32 512 4" "32 32 64 4"; do DSW_OPTIONS=17=1 python tools/time_terms.py $sh 2>/dev/null; DSW_OPTIONS=17=1 DSW_LIB_PATH=/root/repo/deepsphere-weather_b200/libdsw_prev.so python tools/time_terms.py $sh 2>/dev/null; done
