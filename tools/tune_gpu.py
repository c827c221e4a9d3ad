#!/usr/bin/env python
"""GPU-in-the-loop refinement of a pp_interact2 statement order (development tool).

Below ~1 % the register-file model of tools/sass_rf_model.py no longer ranks orders correctly, so the last
step is measured: `gen` writes a population of mutated orders as post-processed cubins under kb_variants/, one
`gpurun -- bash scripts/gpu_cubins.sh` call times them all with kbench (the current order is in every population as the
reference), and `pick` reads gpurun_out/cubins.txt and reports / adopts the winner.

  O3D_TUNE_VARIANT=uni python tools/tune_gpu.py gen  <base.inc> <count> <seed> [max-model-regression]
  python tools/tune_gpu.py pick                      -> prints the ranking; writes kb_variants/best.inc
"""
import json
import os
import random
import re
import shutil
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("O3D_TUNE_PATCHED", "1")
os.environ.setdefault("O3D_TUNE_WORK", "/tmp/kvgpu")
import tune_order as T  # noqa: E402

POP = os.path.join(ROOT, os.environ.get("O3D_TUNE_POP", "kb_variants"))       # population directory (relative to the repo)


def build(args):
    k, order, swaps = args
    r = T.score((k, order, swaps, ["-lineinfo"]))
    if r is None or r[5]:
        return None
    return r, os.path.join(T.WORK, f"one_{k}_p.cubin"), os.path.join(T.WORK, f"body_{k}.inc")


def gen(base, count, seed, slack):
    os.makedirs(T.WORK, exist_ok=True)
    os.makedirs(POP, exist_ok=True)
    for f in os.listdir(POP):
        if f.endswith((".cubin", ".inc")):
            os.remove(os.path.join(POP, f))
    rng = random.Random(seed)
    order, swaps = T.parse_body(base)
    cands = [(0, order, swaps)]
    seen = {(tuple(order), swaps)}
    while len(cands) < count:
        m = T.mutate(rng, order, swaps)
        for _ in range(int(os.environ.get("O3D_TUNE_WIDE", "1")) - 1):   # wider steps: several rounds of 1-3 edits per candidate
            m = T.mutate(rng, *m)
        key = (tuple(m[0]), m[1])
        if key not in seen:
            seen.add(key)
            cands.append((len(cands), m[0], m[1]))
    with ProcessPoolExecutor(int(os.environ.get("O3D_TUNE_JOBS", "6"))) as ex:
        res = list(ex.map(build, cands))
    base_score = res[0][0][0]
    index = {}
    for (k, _, _), r in zip(cands, res):
        if r is None or r[0][0] > base_score + slack:
            continue                                  # keep candidates the model does not rate clearly worse
        name = f"cand_{k:03d}"
        shutil.copy(r[1], os.path.join(POP, name + ".cubin"))
        shutil.copy(r[2], os.path.join(POP, name + ".inc"))
        index[name] = {"model": r[0][0], "three_reads": r[0][2], "regs": r[0][4]}
    json.dump(index, open(os.path.join(POP, "index.json"), "w"), indent=1)
    print(f"{len(index)} candidates under {POP} (cand_000 = the base order, model {base_score:.3f} cycles per body)")


def pick():
    index = json.load(open(os.path.join(POP, "index.json")))
    rows = []
    tag = os.path.relpath(POP, ROOT)
    for l in open(os.path.join(ROOT, "gpurun_out", os.environ.get("O3D_TUNE_TIMES", "cubins.txt"))):
        m = re.match(r"cubin " + re.escape(tag) + r"/(cand_\d+)\.cubin\s+(\w+).*?([\d.]+) ms.*fnv (\w+)", l)
        if m:
            rows.append((float(m.group(3)), m.group(1), m.group(4)))
    rows.sort()
    base = next(t for t, n, _ in rows if n == "cand_000")
    hashes = {h for _, _, h in rows}
    print(f"base {base:.3f} ms; {len(rows)} timed; output hashes: {len(hashes)} distinct (must be 1)")
    for t, n, h in rows[:8]:
        print(f"  {n}  {t:.3f} ms  {100 * (t / base - 1):+.2f} %   model {index[n]['model']:.3f}  three-reads {index[n]['three_reads']}")
    if len(hashes) == 1:
        shutil.copy(os.path.join(POP, rows[0][1] + ".inc"), os.path.join(POP, "best.inc"))
        print("best ->", os.path.join(POP, "best.inc"))


if __name__ == "__main__":
    if sys.argv[1] == "gen":
        gen(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]) if len(sys.argv) > 5 else 0.6)
    else:
        pick()
