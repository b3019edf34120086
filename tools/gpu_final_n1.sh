# round-end record on one B200: default bench line, reference (CPU oracle) arm, launch list, ncu --set full of the pair stage
mkdir -p gpurun_out
(time timeout 900 python bench.py) > gpurun_out/r2_bench_n1.log 2>&1
grep '^{' gpurun_out/r2_bench_n1.log | tail -1 > gpurun_out/r2_bench_n1.json
(time timeout 900 python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/r2_bench_reference.log 2>&1
grep '^{' gpurun_out/r2_bench_reference.log | tail -1 > gpurun_out/r2_bench_reference.json
B="python bench.py --no-cpu --no-parity --derep off"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv $B --steps 2 --warmup 1 > gpurun_out/r2_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'anchor_kernel|chain_kernel|finalize_warp_kernel|cand_pack_kernel|task_setup_kernel|screen_runs_kernel' -s 12 -c 6 -o gpurun_out/r2_final $B --steps 1 --warmup 1 > gpurun_out/r2_final_ncu.log 2>&1
for k in 64 128; do SKB_SEARCH_BATCH=$k timeout 600 python bench.py --workload config4 --no-cpu --derep off --steps 1 --warmup 1 2>/dev/null | grep '^{' | tail -1 > gpurun_out/config4_b$k.json; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print("N=1 value %.1f M (%.2f ms) e2e %.1f M (%.1f ms) derep %.1f s parity %s" % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d.get('derep_wall_s',-1), d['parity_sample']))
r=json.load(open('gpurun_out/r2_bench_reference.json'))
print("reference value %.1f k (%.1f s/step) cores %s" % (r['value']/1e3, r['ms_per_step']/1e3, r['cpu_baseline']['cores']))
for k in (64,128):
    c=json.load(open('gpurun_out/config4_b%d.json'%k)); print("config4 batch", k, "%.1f M pairs/s" % (c['value']/1e6), c['config']['search_batching'])
PY
