L=anqs_quantum_chemistry_b200/libanqs_b200.so
cp $L /tmp/orig.so
for v in A B; do
  cp scripts/_alt/lib_$v.so $L
  echo "== $v"; timeout 300 python scripts/batch_reduce_time.py 87000; timeout 300 python scripts/batch_reduce_time.py 1048576
done
cp /tmp/orig.so $L
timeout 900 python -m pytest tests/test_gpu_anqs.py tests/test_gpu_nade.py tests/test_gpu_transformer.py -x -q 2>&1 | tail -3
