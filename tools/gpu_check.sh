# quick GPU check of a kernel change: parity tests (both diagonal widths), then a traced bench of config3
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/check_tests.log
(SKB_WIDE_DIAG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3) >> gpurun_out/check_tests.log
SKB_TRACE=1 timeout 400 python bench.py --no-cpu --derep off --steps 3 --warmup 3 > gpurun_out/check_bench.log 2> gpurun_out/check_trace.log
cat gpurun_out/check_tests.log
grep "skb trace" gpurun_out/check_trace.log | tail -9
python - <<'PY'
import json
l=[x for x in open('gpurun_out/check_bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print("value %.1f M  ms/step %.2f  e2e %.1f M (%.1f ms)  ms_ani %.2f  anchor %.3f ms  parity %s sha %s" % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['config']['ms_ani'], d['roofline']['ms_per_launch'], d.get('parity_sample',{}).get('mismatch'), d['edges_sha256'][:12]))
PY
