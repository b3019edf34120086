# A/B of prebuilt library variants (tools/build_variants.sh): one short config3 bench per variant
mkdir -p gpurun_out
for so in "$@"; do
  name=$(basename $so .so)
  SKB_LIB=$PWD/$so timeout 300 python bench.py --no-cpu --no-parity --derep off --steps 3 --warmup 2 > gpurun_out/ab_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
try:
    l=[x for x in open('gpurun_out/ab_%s.log'%sys.argv[1]) if x.startswith('{')][-1]
    d=json.loads(l)
    print("%-10s value %.1f M  ms/step %.2f  ms_ani %.2f  anchor %.3f ms/launch  e2e %.1f ms  sha %s" % (sys.argv[1], d['value']/1e6, d['ms_per_step'], d['config']['ms_ani'], d['roofline']['ms_per_launch'], d['e2e']['ms_per_step'], d['edges_sha256'][:12]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
