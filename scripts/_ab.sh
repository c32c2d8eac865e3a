L=anqs_quantum_chemistry_b200/libanqs_b200.so
cp $L /tmp/orig.so
for v in A B A B; do
  cp scripts/_alt/lib_$v.so $L
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'], d['clocks'])"
done
cp /tmp/orig.so $L
