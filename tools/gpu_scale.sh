# multi-GPU check: tools/gpu_scale.sh N [extra bench args]   (run under gpurun --gpus N)
N=$1; shift
mkdir -p gpurun_out
if [ "$N" = "2" ]; then (timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3) > gpurun_out/scale_tests_n$N.log; cat gpurun_out/scale_tests_n$N.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/scale_n$N.log 2>&1
grep '^{' gpurun_out/scale_n$N.log | tail -1 > gpurun_out/r2_bench_n$N.json
python - $N <<'PY'
import json,sys
n=sys.argv[1]
ls=[x for x in open('gpurun_out/scale_n%s.log'%n) if x.startswith('{')]
if not ls: print(open('gpurun_out/scale_n%s.log'%n).read()[-3000:]); sys.exit(0)
d=json.loads(ls[-1])
print("N=%s value %.1f M (%.2f ms)  e2e %.1f M (%.1f ms: upload %.1f repl %.1f index %.1f)  ms_ani %.2f ms_screen %.2f sha %s" % (n, d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['ms_upload_sketch'], d['e2e']['ms_replicate'], d['e2e']['ms_index'], d['config']['ms_ani'], d['config']['ms_screen'], d['edges_sha256'][:12]))
PY
