for cfg in "A:" "B:DSW_FORK=0" "C:DSW_OPTIONS=2=1024" "D:DSW_FUSED_SKIPS=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python tools/step_kineto.py 3 2>/dev/null | grep -E "^#|mix_tma_kernel<true>|CUDAFunctor_a" | sed "s/^/$name /"
done
