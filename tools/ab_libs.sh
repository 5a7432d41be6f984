# A/B of library builds under bench.py's per-kernel timing (L2 flushed between reps)
for v in orig head new; do
  SLDM_LIB_PATH=$PWD/build/ab/$v.so python bench.py --steps 10 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'step %.3f ms' % d['ms_per_step'], ' '.join('%s=%.4f' % (k[:14], v['ms']) for k, v in d['kernels'].items()))"
done
