timeout 900 python -m pytest tests/test_encode_gpu.py tests/test_stripes_gpu.py tests/test_encode_fuzz_gpu.py tests/test_headline_gpu.py tests/test_surface_gpu.py -m gpu -x -q 2>&1 | tail -8
echo "== v1 kernel"
B2V_SLICE_KERNEL=v1 timeout 900 python -m pytest tests/test_encode_gpu.py tests/test_stripes_gpu.py -m gpu -x -q 2>&1 | tail -4
for v in "X=0" "B2V_SLICE_ROWS=1" "B2V_SLICE_KERNEL=v1" "B2V_SLICE_ROWS=135"; do echo "== $v"; env $v timeout 250 python tools/instep.py 384 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['plain']['fps']), {k:round(v,1) for k,v in d['timing_all'].items()})"; done
