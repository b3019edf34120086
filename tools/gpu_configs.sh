# the other BASELINE.json workloads on one GPU: config2 (1,000 genomes), config3r (real-sequence-like), config4 (search path)
mkdir -p gpurun_out
for w in config2 config3r; do
  timeout 600 python bench.py --workload $w --no-cpu --derep off --steps 5 --warmup 3 > gpurun_out/r2_bench_${w}_n1.log 2>&1
  grep '^{' gpurun_out/r2_bench_${w}_n1.log | tail -1 > gpurun_out/r2_bench_${w}_n1.json
done
timeout 900 python bench.py --workload config4 --no-cpu --derep off --steps 2 --warmup 1 > gpurun_out/r2_bench_config4_n1.log 2>&1
grep '^{' gpurun_out/r2_bench_config4_n1.log | tail -1 > gpurun_out/r2_bench_config4_n1.json
python - <<'PY'
import json
for w in ("config2","config3r","config4"):
    try:
        d=json.load(open('gpurun_out/r2_bench_%s_n1.json'%w))
        print(w, "value %.2f M %s  ms/step %.2f  e2e %.2f M (%.1f ms)" % (d['value']/1e6, d['unit'], d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step']), {k:d['config'].get(k) for k in ('pairs','edges','representatives','searches_per_s','ms_per_search')}, d.get('parity_sample'))
    except Exception as e:
        print(w, "FAILED", e); print(open('gpurun_out/r2_bench_%s_n1.log'%w).read()[-1500:])
PY
