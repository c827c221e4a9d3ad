#!/usr/bin/env python
"""Opcode histogram and SASS listing of a kernel's hot loops (evidence tool, no GPU needed).

usage: sass_hotloop.py <cubin|lib.so> <kernel-name-substring> [--list]   > profiles/<name>.txt
Finds every innermost loop that holds a MUFU instruction (the interaction loops of pp2_kernel / ppc_kernel), prints per loop
the instruction count by opcode for one trip, and with --list the instructions of the first such loop.
"""
import collections
import re
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
import sass_rf_model as M  # noqa: E402


def main():
    path, name = sys.argv[1], sys.argv[2]
    ins = M.kernel_sass(path, name)
    whole = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
    print(f"# {path}: kernel *{name}*: {len(ins)} instructions")
    print("# whole kernel, by opcode: " + ", ".join(f"{k} {v}" for k, v in whole.most_common(24)))
    loops = M.hot_loops(ins)
    for j, i in loops:
        body = [t for _, t in ins[j:i + 1]]
        c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
        n, three, tot, other = M.model(body)
        print(f"\nloop {ins[j][0]:#06x}..{ins[i][0]:#06x}: {len(body)} instructions per trip; packed FP32 {n} ({three} need a third register-file cycle), "
              f"modelled FMA-pipe cycles {tot} (ideal {2 * n})")
        print("  " + ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])))
        reuse = sum(t.count(".reuse") for t in body)
        print(f"  operand-reuse flags: {reuse}")
    if "--list" in sys.argv and loops:
        j, i = loops[0]
        print(f"\n# first loop, {i - j + 1} instructions:")
        for a, t in ins[j:i + 1]:
            print(f"  /*{a:04x}*/  {t} ;")


if __name__ == "__main__":
    main()
