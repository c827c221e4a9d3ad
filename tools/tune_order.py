#!/usr/bin/env python
"""Offline search over the statement order of the packed velocity+gradient interaction (development tool, no GPU).

The inner loop of pp2_kernel is bound by the FP32 pipe, and what separates one legal ordering of its ~35 packed
instructions from another is how many of them need a third register-file cycle (tools/sass_rf_model.py; the model has
tracked B200 measurements to ~1 % three times: profiles/r01_variants_*.txt). ptxas keeps much of the source order of
independent instructions, so this script writes the body of pp_interact2<GRAD=true, UNI=true> as a random topological
order of its dataflow graph (with random operand swaps of the commutative products), compiles only that kernel
(~1.3 s) and scores the SASS with the model. The best orders are then timed on the GPU (kbench) before one is adopted.

usage: tune_order.py [trials] [seed] [jobs]      -> prints the best bodies; writes /tmp/kv/best_<k>.inc
"""
import os
import random
import re
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_rf_model as M  # noqa: E402

WORK = os.environ.get("O3D_TUNE_WORK", "/tmp/kv")

# (result, op, a, b, c) - c None for 2-operand ops. Names starting with '-' are negated operands.
# acc updates read and write acc[k]. Commutative operand pairs (a, b) may be swapped.
STMTS = [
    ("dx", "add", "tx", "sx", None), ("dy", "add", "ty", "sy", None), ("dz", "add", "tz", "sz", None),
    ("a1", "fma", "dz", "dz", "r2"), ("a2", "fma", "dy", "dy", "a1"), ("d2", "fma", "dx", "dx", "a2"),
    ("rs", "rsq", "d2", None, None),
    ("rs2", "mul", "rs", "rs", None), ("rs3", "mul", "rs2", "rs", None), ("dn5", "mul", "rs3", "rs2", None),
    ("r3", "fma", "tk", "dn5", "rs3"),
    ("t1", "mul", "dy", "wz", None), ("t2", "mul", "dx", "wz", None), ("t3", "mul", "dx", "wy", None),
    ("cx", "fma", "dz", "wy", "-t1"), ("cy", "fma", "-dz", "wx", "t2"), ("cz", "fma", "dy", "wx", "-t3"),
    ("acc[12]", "fma", "r3", "wx", "acc[12]"), ("acc[13]", "fma", "r3", "wy", "acc[13]"), ("acc[14]", "fma", "r3", "wz", "acc[14]"),
    ("acc[0]", "fma", "r3", "cx", "acc[0]"), ("acc[1]", "fma", "r3", "cy", "acc[1]"), ("acc[2]", "fma", "r3", "cz", "acc[2]"),
    ("w", "fma", "tk5", "rs2", "m3"), ("bbb", "mul", "dn5", "w", None),
    ("bx", "mul", "bbb", "cx", None), ("by", "mul", "bbb", "cy", None), ("bz", "mul", "bbb", "cz", None),
    ("acc[3]", "fma", "dx", "bx", "acc[3]"), ("acc[4]", "fma", "dx", "by", "acc[4]"), ("acc[5]", "fma", "dx", "bz", "acc[5]"),
    ("acc[6]", "fma", "dy", "bx", "acc[6]"), ("acc[7]", "fma", "dy", "by", "acc[7]"), ("acc[8]", "fma", "dy", "bz", "acc[8]"),
    ("acc[9]", "fma", "dz", "bx", "acc[9]"), ("acc[10]", "fma", "dz", "by", "acc[10]"),
]
INPUTS = {"tx", "ty", "tz", "sx", "sy", "sz", "r2", "tk", "tk5", "m3", "wx", "wy", "wz"} | {f"acc[{k}]" for k in range(15)}

# O3D_TUNE_VARIANT selects which instantiation of pp_interact2 is searched (default: velocity+gradient, uniform radii):
#   gen    velocity+gradient, per-particle radii (38 statements)      -> -DO3D_PP_BODY_FILE_GEN,    pp2_kernel<2,true,128>
#   vel    velocity only, uniform radii (19)                          -> -DO3D_PP_BODY_FILE_VEL,    pp2_kernel<4,false,128>
#   velgen velocity only, per-particle radii (21)                     -> -DO3D_PP_BODY_FILE_VELGEN, pp2_kernel<4,false,128>
VARIANT = os.environ.get("O3D_TUNE_VARIANT", "uni")
_GEOM = [("dx", "add", "tx", "sx", None), ("dy", "add", "ty", "sy", None), ("dz", "add", "tz", "sz", None)]
_D2 = [("a1", "fma", "dz", "dz", "r2"), ("a2", "fma", "dy", "dy", "a1"), ("d2", "fma", "dx", "dx", "a2"), ("rs", "rsq", "d2", None, None)]
_CROSS = [("t1", "mul", "dy", "wz", None), ("t2", "mul", "dx", "wz", None), ("t3", "mul", "dx", "wy", None),
          ("cx", "fma", "dz", "wy", "-t1"), ("cy", "fma", "-dz", "wx", "t2"), ("cz", "fma", "dy", "wx", "-t3")]
_VEL = [("acc[0]", "fma", "r3", "cx", "acc[0]"), ("acc[1]", "fma", "r3", "cy", "acc[1]"), ("acc[2]", "fma", "r3", "cz", "acc[2]")]
_ANTI = [("acc[12]", "fma", "r3", "wx", "acc[12]"), ("acc[13]", "fma", "r3", "wy", "acc[13]"), ("acc[14]", "fma", "r3", "wz", "acc[14]")]
_GRAD = [("bx", "mul", "bbb", "cx", None), ("by", "mul", "bbb", "cy", None), ("bz", "mul", "bbb", "cz", None),
         ("acc[3]", "fma", "dx", "bx", "acc[3]"), ("acc[4]", "fma", "dx", "by", "acc[4]"), ("acc[5]", "fma", "dx", "bz", "acc[5]"),
         ("acc[6]", "fma", "dy", "bx", "acc[6]"), ("acc[7]", "fma", "dy", "by", "acc[7]"), ("acc[8]", "fma", "dy", "bz", "acc[8]"),
         ("acc[9]", "fma", "dz", "bx", "acc[9]"), ("acc[10]", "fma", "dz", "by", "acc[10]")]
# per-particle radii: r2 = sr^2 + tr^2 per pair, top = d2 + 1.5 r2, dn5 = rs^4 rs, r3 = top dn5, bbb = dn5 (2 - 5 top rs^2)
_CORE_GEN = [("r2", "add", "tr2", "sr2", None), ("rs2", "mul", "rs", "rs", None), ("top", "fma", "c15", "r2", "d2"),
             ("rs4", "mul", "rs2", "rs2", None), ("dn5", "mul", "rs4", "rs", None), ("r3", "mul", "top", "dn5", None)]
_BBB_GEN = [("tq", "mul", "top", "rs2", None), ("w", "fma", "m5", "tq", "p2"), ("bbb", "mul", "dn5", "w", None)]
_CORE_UNI = [("rs2", "mul", "rs", "rs", None), ("rs3", "mul", "rs2", "rs", None), ("dn5", "mul", "rs3", "rs2", None), ("r3", "fma", "tk", "dn5", "rs3")]
if VARIANT == "gen":
    STMTS = _GEOM + [_CORE_GEN[0]] + _D2 + _CORE_GEN[1:] + _CROSS + _ANTI + _VEL + _BBB_GEN + _GRAD
elif VARIANT == "vel":
    STMTS = _GEOM + _D2 + _CORE_UNI + _CROSS + _VEL
elif VARIANT == "velgen":
    STMTS = _GEOM + [_CORE_GEN[0]] + _D2 + _CORE_GEN[1:] + _CROSS + _VEL
# the alternate core functions of src/CoreFunc.h (csrc/biot_pp_cores.cuh: ppc_kernel<CORE, T, GRAD, 128>):
#   rm / rmvel   Rosenhead-Moore, velocity+gradient (34 statements) / velocity only (18)
#   v2 / v2vel   Vatistas n=2 (34 / 18; two MUFU per lane: rsqrt and sqrt)
_ST = [("st", "add", "tt", "sl", None)]
_D2_RM = [("a1", "fma", "dz", "dz", "st"), ("a2", "fma", "dy", "dy", "a1"), ("d2", "fma", "dx", "dx", "a2"), ("rs", "rsq", "d2", None, None)]
_CORE_RM = [("rs2", "mul", "rs", "rs", None), ("r3", "mul", "rs2", "rs", None)]
_BBB_RM = [("w", "mul", "m3", "rs2", None), ("bbb", "mul", "w", "r3", None)]
_D2_V2 = [("b1", "mul", "dz", "dz", None), ("b2", "fma", "dy", "dy", "b1"), ("dsq", "fma", "dx", "dx", "b2"), ("den", "fma", "dsq", "dsq", "st"),
          ("rq", "rsq", "den", None, None), ("sq", "sqrt", "rq", None, None)]
_CORE_V2 = [("r3", "mul", "rq", "sq", None)]
_BBB_V2 = [("w", "mul", "m3", "rq", None), ("bbb", "mul", "w", "r3", None)]
if VARIANT == "rm":
    STMTS = _GEOM + _ST + _D2_RM + _CORE_RM + _CROSS + _ANTI + _VEL + _BBB_RM + _GRAD
elif VARIANT == "rmvel":
    STMTS = _GEOM + _ST + _D2_RM + _CORE_RM + _CROSS + _VEL
elif VARIANT == "v2":
    STMTS = _GEOM + _ST + _D2_V2 + _CORE_V2 + _CROSS + _ANTI + _VEL + _BBB_V2 + _GRAD
elif VARIANT == "v2vel":
    STMTS = _GEOM + _ST + _D2_V2 + _CORE_V2 + _CROSS + _VEL
PPC = VARIANT in ("rm", "rmvel", "v2", "v2vel")
HEADER = "biot_pp_cores.cuh" if PPC else "biot_pp.cuh"
SASS_NAME = "ppc_kernel" if PPC else "pp2_kernel"
MUFU_PER_BODY = 4.0 if VARIANT in ("v2", "v2vel") else 2.0     # MUFU instructions per (target, source pair): two lanes x (rsqrt [+ sqrt])
if PPC:
    KERNEL = {"rm": "ppc_kernel<1, 2, true, 384>", "rmvel": "ppc_kernel<1, 4, false, 384>",
              "v2": "ppc_kernel<3, 2, true, 384>", "v2vel": "ppc_kernel<3, 4, false, 384>"}[VARIANT]
    BODY_MACRO = {"rm": "O3D_PPC_BODY_RM_GRAD", "rmvel": "O3D_PPC_BODY_RM_VEL", "v2": "O3D_PPC_BODY_V2_GRAD", "v2vel": "O3D_PPC_BODY_V2_VEL"}[VARIANT]
else:
    KERNEL = "pp2_kernel<2, true, 384>" if VARIANT in ("uni", "gen") else "pp2_kernel<4, false, 384>"
    BODY_MACRO = {"uni": "O3D_PP_BODY_FILE", "gen": "O3D_PP_BODY_FILE_GEN", "vel": "O3D_PP_BODY_FILE_VEL", "velgen": "O3D_PP_BODY_FILE_VELGEN"}[VARIANT]
PER_BODY = len([st for st in STMTS if st[1] not in ("rsq", "sqrt")])        # packed instructions per (target, source pair): picks the loop

# O3D_TUNE_JOINT=1: search ONE order over the statements of both register-blocked targets (T = 2) - the source operands
# (wx wy wz sx sy sz and the constants) are shared between the two, so operand-reuse chains can span the targets.
JOINT = os.environ.get("O3D_TUNE_JOINT", "0") == "1"
PATCHED = os.environ.get("O3D_TUNE_PATCHED", "0") == "1"     # score the SASS after the reuse-chain post-pass
T0 = float(os.environ.get("O3D_TUNE_TEMP", "0.6"))
EXTRA = os.environ.get("O3D_TUNE_EXTRA", "").split()         # extra nvcc flags for every candidate, e.g. -DO3D_PP_UNROLL_GRAD=1
SHARED = {"sx", "sy", "sz", "r2", "tk", "tk5", "m3", "wx", "wy", "wz"}


def _rename(name, t):
    if name is None:
        return None
    neg = name.startswith("-")
    n = name.lstrip("-")
    if n in SHARED:
        out = n
    elif n in ("tx", "ty", "tz"):
        out = f"{n}[{t}]"
    elif n.startswith("acc["):
        out = f"acc[{t}]{n[3:]}"
    else:
        out = n + "AB"[t]
    return ("-" if neg else "") + out


if JOINT:
    STMTS = [(_rename(r, t), op, _rename(a, t), _rename(b, t), _rename(c, t)) for t in (0, 1) for (r, op, a, b, c) in STMTS]

PRELUDE = """    const float2 sx = f2(q0.x, q0.y), sy = f2(q0.z, q0.w), sz = f2(q1.x, q1.y);
    const float2 wx = f2(q2.x, q2.y), wy = f2(q2.z, q2.w), wz = f2(q3.x, q3.y);
    const float2 r2 = tr2, m3 = f2(-3.0f, -3.0f);
"""
if VARIANT in ("gen", "velgen"):
    PRELUDE = PRELUDE.replace("const float2 r2 = tr2, m3 = f2(-3.0f, -3.0f);",
                              "const float2 sr2 = f2(q1.z, q1.w), c15 = f2(1.5f, 1.5f), m5 = f2(-5.0f, -5.0f), p2 = f2(2.0f, 2.0f);")
if PPC:
    PRELUDE = PRELUDE.replace("const float2 r2 = tr2, m3 = f2(-3.0f, -3.0f);", "const float2 sl = f2(q1.z, q1.w), m3 = f2(-3.0f, -3.0f);")
if JOINT:
    PRELUDE = PRELUDE.replace("r2 = tr2,", "r2 = tr2[0],")


def emit(order, swaps):
    out = [PRELUDE]
    for idx in order:
        res, op, a, b, c = STMTS[idx]
        if idx in swaps and op in ("mul", "fma", "add"):
            a, b = b, a
        def v(x):
            return f"neg2({x[1:]})" if x.startswith("-") else x
        decl = "" if res.startswith("acc[") else "const float2 "
        if op == "add" and res == "st":   # alternate cores: with one radius in the system the pair term arrives precomputed in tt
            out.append(f"    {decl}{res} = UNI ? tt : __fadd2_rn({v(a)}, {v(b)});\n")
        elif op == "add":
            out.append(f"    {decl}{res} = __fadd2_rn({v(a)}, {v(b)});\n")
        elif op == "mul":
            out.append(f"    {decl}{res} = __fmul2_rn({v(a)}, {v(b)});\n")
        elif op == "fma":
            out.append(f"    {decl}{res} = __ffma2_rn({v(a)}, {v(b)}, {v(c)});\n")
        elif op == "rsq":
            out.append(f"    {decl}{res} = f2(rsqrt_approx({a}.x), rsqrt_approx({a}.y));\n")
        elif op == "sqrt":
            out.append(f"    {decl}{res} = f2(sqrt_approx({a}.x), sqrt_approx({a}.y));\n")
    return "".join(out)


def deps(idx):
    _, _, a, b, c = STMTS[idx]
    names = {x.lstrip("-") for x in (a, b, c) if x}
    return {k for k, st in enumerate(STMTS) if st[0] in names and not st[0].startswith("acc[")} - {idx}


DEPS = [deps(i) for i in range(len(STMTS))]


def operands(idx):
    _, _, a, b, c = STMTS[idx]
    return {x.lstrip("-") for x in (a, b) if x}


def random_order(rng, chain_bias):
    done, order = set(), []
    prev = None
    while len(order) < len(STMTS):
        ready = [i for i in range(len(STMTS)) if i not in done and DEPS[i] <= done]
        pick = None
        if prev is not None and rng.random() < chain_bias:
            share = [i for i in ready if operands(i) & operands(prev)]
            if share:
                pick = rng.choice(share)
        if pick is None:
            pick = rng.choice(ready)
        order.append(pick)
        done.add(pick)
        prev = pick
    return order


def score(args):
    k, order, swaps, extra = args
    body = os.path.join(WORK, f"body_{k}.inc")
    with open(body, "w") as f:
        f.write(emit(order, swaps))
    cu = os.path.join(WORK, f"one_{k}.cu")
    with open(cu, "w") as f:
        f.write('#include "' + HEADER + '"\nusing namespace o3d;\ntemplate __global__ void o3d::' + KERNEL + '(const PPArgs);\n')
    cubin = os.path.join(WORK, f"one_{k}.cubin")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-I" + os.path.join(ROOT, "omega3d_b200", "csrc"),
           "-DO3D_PP_POW=2", f'-D{BODY_MACRO}="{body}"', "-Xptxas", "-v", "-cubin", cu, "-o", cubin] + extra + EXTRA
    if JOINT:
        cmd[cmd.index("-DO3D_PP_POW=2") + 1] = "-DO3D_PP_JOINT=1"
        cmd.append(f'-DO3D_PP_JOINT_FILE="{body}"')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return None
    if PATCHED:                                   # score what tools/sass_patch.py makes of it
        pc = os.path.join(WORK, f"one_{k}_p.cubin")
        if subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_patch.py"), cubin, pc, SASS_NAME],
                          capture_output=True, text=True).returncode != 0:
            return None
        cubin = pc
    regs = int(re.search(r"Used (\d+) registers", r.stderr).group(1))
    spill = "0 bytes spill stores" not in r.stderr
    ins = M.kernel_sass(cubin, SASS_NAME)
    # The kernel holds a uniform-radius and a per-particle-radius variant of the loop, each in TWO copies (one per ring
    # buffer: pp2_tile<0>, pp2_tile<1>) that ptxas allocates registers for independently; a walk alternates the two copies,
    # so the score is their mean.
    found = []
    for j, i in M.hot_loops(ins):
        bodyins = [t for _, t in ins[j:i + 1]]
        n, three, tot, other = M.model(bodyins)
        nm = sum(1 for t in bodyins if t.startswith("MUFU"))
        if nm and abs(n / (nm / MUFU_PER_BODY) - PER_BODY) < 0.5:
            found.append((n, three, tot, other, nm))
    if not found:
        return None
    k = len(found)
    n, three, cyc, nm = (sum(f[0] for f in found) / k, sum(f[1] for f in found) / k, sum(f[2] + f[3] for f in found) / k,
                         sum(f[4] for f in found) / k)
    per_pair = cyc / (nm / MUFU_PER_BODY)                  # modelled cycles per (target, source pair)
    return per_pair, n, three, cyc, regs, spill


def legal(order):
    pos = {s: k for k, s in enumerate(order)}
    return all(pos[d] < pos[i] for i in order for d in DEPS[i])


def slot_operands(idx, swaps):
    """(slot a, slot b) operand names of a statement as emitted (after its optional swap), negation stripped."""
    _, op, a, b, _ = STMTS[idx]
    if op == "rsq":
        return (None, None)
    if idx in swaps:
        a, b = b, a
    return (a.lstrip("-") if a else None, b.lstrip("-") if b else None)


def mutate(rng, order, swaps):
    """One to three random edits: flip an operand swap, move a statement to a random legal place, or ("chain" move) put a
    statement next to one that reads the same operand in the same slot - the shape an operand-reuse chain has."""
    order, swaps = list(order), set(swaps)
    for _ in range(rng.choice([1, 1, 2, 3])):
        r = rng.random()
        if r < 0.2:
            swaps ^= {rng.randrange(len(STMTS))}
            continue
        if r < 0.6:
            for _ in range(50):
                i = rng.randrange(len(order))
                oa = slot_operands(order[i], swaps)
                mates = [k for k in range(len(order)) if k != i and any(x is not None and x == y for x, y in zip(oa, slot_operands(order[k], swaps)))]
                if not mates:
                    continue
                k = rng.choice(mates)
                cand = list(order)
                st = cand.pop(k)
                pos = cand.index(order[i]) + rng.choice([0, 1])
                cand.insert(pos, st)
                if legal(cand):
                    order = cand
                    break
            continue
        for _ in range(50):
            i = rng.randrange(len(order))
            j = rng.randrange(len(order))
            cand = list(order)
            cand.insert(j, cand.pop(i))
            if legal(cand):
                order = cand
                break
    return order, frozenset(swaps)


def climb(rounds, seed, jobs):
    """Hill climbing from the hand-written order: `jobs` mutants per round, keep the best if it improves."""
    rng = random.Random(seed)
    best = (list(range(len(STMTS))), frozenset())
    best_r = score((0, best[0], best[1], []))
    print("start", best_r)
    for rnd in range(rounds):
        muts = [mutate(rng, *best) for _ in range(jobs)]
        with ProcessPoolExecutor(jobs) as ex:
            res = list(ex.map(score, [(k + 1, m[0], m[1], []) for k, m in enumerate(muts)]))
        ok = [(r, m) for r, m in zip(res, muts) if r is not None and not r[5]]
        if not ok:
            continue
        r, m = min(ok, key=lambda t: t[0][0])
        if r[0] <= best_r[0]:
            if r[0] < best_r[0]:
                print(f"round {rnd}: {r}", flush=True)
            best, best_r = m, r
    with open(os.path.join(WORK, f"climb_{seed}.inc"), "w") as f:
        f.write(emit(best[0], best[1]))
    print("final", best_r, "->", os.path.join(WORK, f"climb_{seed}.inc"))


def parse_body(path):
    """Recover (order, swaps) from a body file written by emit()."""
    order, swaps = [], set()
    for line in open(path):
        m = re.match(r"\s+(?:const float2 )?([\w\[\]]+) = (?:UNI \? tt : )?(__f\w+2_rn|f2)\((.*)\);", line)
        if not m or line.lstrip().startswith(("const float2 sx =", "const float2 wx =", "const float2 r2 = tr2", "const float2 sr2 = f2(q1", "const float2 sl = f2(q1")):
            continue
        res = m.group(1)
        idx = next(k for k, st in enumerate(STMTS) if st[0] == res)
        order.append(idx)
        if STMTS[idx][1] in ("mul", "fma", "add"):
            first = m.group(3).split(",")[0].strip()
            a = STMTS[idx][2]
            want = f"neg2({a[1:]})" if a.startswith("-") else a
            if first != want:
                swaps.add(idx)
    assert sorted(order) == list(range(len(STMTS))), "body file does not hold every statement once"
    return order, frozenset(swaps)


def anneal(minutes, seed, jobs, start_file=None):
    """Simulated annealing (parallel tempering-lite): `jobs` mutants of the current state per round, Metropolis
    acceptance on the best of them; remembers the best state ever seen."""
    import math
    import time
    rng = random.Random(seed)
    cur = parse_body(start_file) if start_file else (list(range(len(STMTS))), frozenset())
    cur_r = score((0, cur[0], cur[1], []))
    best, best_r = cur, cur_r
    print("start", cur_r, flush=True)
    t_end = time.time() + 60 * minutes
    rnd = 0
    while time.time() < t_end:
        frac = max(0.0, (t_end - time.time()) / (60 * minutes))
        temp = T0 * (0.08 + frac)                    # in cycles per (target, source pair); O3D_TUNE_TEMP sets T0
        muts = [mutate(rng, *cur) for _ in range(jobs)]
        with ProcessPoolExecutor(jobs) as ex:
            res = list(ex.map(score, [(k + 1, m[0], m[1], []) for k, m in enumerate(muts)]))
        ok = [(r, m) for r, m in zip(res, muts) if r is not None and not r[5]]
        rnd += 1
        if not ok:
            continue
        r, m = min(ok, key=lambda t: t[0][0])
        if r[0] <= cur_r[0] or rng.random() < math.exp(-(r[0] - cur_r[0]) / temp):
            cur, cur_r = m, r
        if r[0] < best_r[0]:
            best, best_r = m, r
            print(f"round {rnd}: {r}", flush=True)
            with open(os.path.join(WORK, f"anneal_{seed}.inc"), "w") as f:
                f.write(emit(best[0], best[1]))
    print("final", best_r, flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "emit":      # write the formula-order body of the variant to a file
        with open(sys.argv[2], "w") as f:
            f.write(emit(list(range(len(STMTS))), frozenset()))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "anneal":
        return anneal(float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 8,
                      sys.argv[5] if len(sys.argv) > 5 else None)
    if len(sys.argv) > 1 and sys.argv[1] == "climb":
        return climb(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 8)
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    jobs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    os.makedirs(WORK, exist_ok=True)
    rng = random.Random(seed)
    cands = [(0, list(range(len(STMTS))), frozenset(), [])]          # the hand-written order
    for k in range(1, trials):
        order = random_order(rng, rng.choice([0.0, 0.5, 0.8, 0.95]))
        swaps = frozenset(i for i in range(len(STMTS)) if rng.random() < 0.3)
        cands.append((k, order, swaps, []))
    with ProcessPoolExecutor(jobs) as ex:
        res = list(ex.map(score, cands))
    ranked = sorted((r, c) for r, c in zip(res, cands) if r is not None and not r[5])
    print(f"hand-written order: {res[0]}")
    for rank, (r, c) in enumerate(ranked[:5]):
        print(f"#{rank}: cycles/pair {r[0]:.2f}  fp32 {r[1]}  three-read {r[2]}  cycles/trip {r[3]}  regs {r[4]}   (candidate {c[0]})")
        with open(os.path.join(WORK, f"best_{seed}_{rank}.inc"), "w") as f:
            f.write(emit(c[1], c[2]))
    worst = ranked[-1][0]
    print(f"worst: cycles/pair {worst[0]:.2f}; spread {worst[0] / ranked[0][0][0] - 1:.1%} over {len(ranked)} orders")


if __name__ == "__main__":
    main()
