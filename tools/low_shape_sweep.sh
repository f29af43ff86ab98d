# q34 default schedule with every CTA shape of the rotated final pass (warps x log2 inputs per CTA)
for shape in 8x8 8x10 4x7 2x6 1x5; do
  QCM_LOW_SHAPE=$shape python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$shape', 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel'], 'frac %.4f'%d['roofline']['frac'], d['check']['parity_ok'])"
done
