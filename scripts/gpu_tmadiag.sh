for lib in yolo_somi_b200/libdcnv3_sm100.so scripts/experiments/diag/libdcnv3_td1.so scripts/experiments/diag/libdcnv3_td2.so yolo_somi_b200/libdcnv3_sm100.so; do
  DCNV3_SM100_LIB=$PWD/$lib timeout 150 python bench.py --steps 30 --warmup 5 --no-cpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['passes']
print('$lib fwd %.1f us bwd %.1f us step %.1f us' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3, d['ms_per_step']*1e3))"
done
