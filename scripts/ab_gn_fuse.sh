for v in NMM_NO_GN_FUSE=1 NMM_X=1; do
env $v timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v', round(d['ms_per_step'],3), 'gemm TF/s', round(r['achieved'],1), 'share', round(r['share_of_step'],3), 'launches', r['launches'], {k:round(v['ms_per_step'],3) for k,v in r['other_kernels'].items()}, r.get('measured_over','')[-60:])
"
done
