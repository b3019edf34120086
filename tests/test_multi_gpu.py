"""Two-GPU run of the reference-sharded triangle (needs 2 visible GPUs; skipped otherwise): the edge list gathered on
rank 0 equals the single-GPU one."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, hashlib
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %r)
from skder_b200 import engine, multi, synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
gens = [g for c in range(3) for g in synth.clade_of("tinyr", c)]           # 12 genomes, 3 clades, repeats + indels
mine = [g for i, g in enumerate(gens) if (i // 4) %% world == rank]        # whole clades per rank, as bench.py deals them
order = [i for r in range(world) for i in range(len(gens)) if (i // 4) %% world == r]
eng = engine.Engine(rank)
eng.add([engine.pack_contigs(g) for g in mine])
multi.replicate_sketches(eng, dist, torch, sharded_index=True)
eng.index()
_, st, st_screen = multi.triangle_sharded(eng, dist, torch, 80.0, 15.0, to_host=False)
edges = multi.gather_device_edges(eng, dist, torch, sort=False)
if rank == 0:
    canon = np.array(order)
    rows = sorted((min(canon[a], canon[b]), max(canon[a], canon[b]), "%%.2f" %% ani,
                   "%%.2f" %% (afa if canon[a] < canon[b] else afb), "%%.2f" %% (afb if canon[a] < canon[b] else afa))
                  for a, b, ani, afa, afb in edges.tolist())
    with engine.Engine(0) as one:
        one.add([engine.pack_contigs(g) for g in gens])
        one.index()
        ref, _ = one.triangle(80.0, 15.0)
    want = sorted((a, b, "%%.2f" %% ani, "%%.2f" %% afa, "%%.2f" %% afb) for a, b, ani, afa, afb in ref.tolist())
    print("RESULT", json.dumps({"equal": rows == want, "n": len(rows), "n_want": len(want)}))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_gpu_sharded_triangle_equals_single_gpu(built_lib, tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(29700 + os.getpid() % 200), str(w)], capture_output=True, text=True, timeout=600)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")]
    assert out.returncode == 0 and line, out.stdout[-2000:] + out.stderr[-3000:]
    import json

    r = json.loads(line[0][7:])
    assert r["equal"] and r["n"] == r["n_want"] == 18, r


def test_two_gpu_search_loop_equals_single_gpu(built_lib):
    """config4's multi-GPU design (database sharded per rank, query sketch broadcast, hit ids all-gathered) selects the
    same number of representatives as one GPU holding the whole database."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_gpu_parity import _bench_line

    one = _bench_line(["--workload", "tiny4", "--steps", "1", "--warmup", "0"])
    two = _bench_line(["--workload", "tiny4", "--gpus", "2", "--steps", "1", "--warmup", "0"], nproc=2)
    assert two["n_gpus"] == 2 and two["config"]["representatives"] == one["config"]["representatives"]
    assert two["config"]["reps_sha256"] == one["config"]["reps_sha256"]


def test_two_gpu_bench_checksum_equals_single_gpu(built_lib):
    """bench.py's edge checksum: N = 2 (reference-sharded) and N = 1 produce the same canonical edge list."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_gpu_parity import _bench_line

    a = _bench_line(["--workload", "tinyr", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-parity", "--derep", "off"])
    b = _bench_line(["--workload", "tinyr", "--gpus", "2", "--steps", "1", "--warmup", "1", "--no-cpu", "--no-parity", "--derep", "off"], nproc=2)
    assert a["edges_sha256"] == b["edges_sha256"] and a["config"]["edges"] == b["config"]["edges"] > 0
