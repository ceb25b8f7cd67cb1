"""The C++ host (tools/trueknn) driving N GPUs through tknn_create_multi, checked against the CPU oracle:

    python tools/cli_multi_check.py <n_gpus> [n_points] [k]

Writes the uniform cloud as a raw .f32 point file, runs `tools/trueknn <file> n 3 0 k out --gpus N --mode shard` and
`--mode partition` with binary neighbour output, and compares every row with oracle/knn_oracle.c's kd-tree.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from owlraytracing_b200 import datasets  # noqa: E402

gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
x = datasets.uniform(n, 42)
ref_i, ref_d = O.knn_kdtree(x, k)
ok_all = True
with tempfile.TemporaryDirectory() as tmp:
    pts = os.path.join(tmp, "pts.f32")
    x.tofile(pts)
    for mode in ("shard", "partition"):
        nn = os.path.join(tmp, f"nn_{mode}")
        r = subprocess.run([os.path.join(ROOT, "tools", "trueknn"), pts, str(n), "3", "0", str(k), os.path.join(tmp, "time.txt"),
                            "--neighbours", nn, "--binary", "--json", "--gpus", str(gpus), "--mode", mode], capture_output=True, text=True,
                           timeout=600)
        if r.returncode != 0:
            print(json.dumps({"mode": mode, "rc": r.returncode, "stderr": r.stderr[-400:]}))
            ok_all = False
            continue
        idx = np.fromfile(nn + ".idx.i32", dtype=np.int32).reshape(n, k)
        dist = np.fromfile(nn + ".dist.f32", dtype=np.float32).reshape(n, k)
        ok = bool((idx == ref_i).all() and np.allclose(dist, ref_d, rtol=1e-6, atol=0))
        ok_all = ok_all and ok
        line = json.loads(r.stdout.strip().splitlines()[-1])
        line.update({"equals_oracle": ok, "host": "tools/trueknn (C++) -> tknn_create_multi"})
        print(json.dumps(line), flush=True)
sys.exit(0 if ok_all else 1)
