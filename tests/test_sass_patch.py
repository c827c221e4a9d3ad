"""tools/sass_patch.py (the SASS post-pass behind omega3d_b200/lib/pp2_tuned.cubin) - checked without a GPU:
the patched cubin differs from the compiled one ONLY in operand-reuse bits (58-60) and the yield bit (45) of the upper
control word of packed FP32 instructions, every flag it sets obeys the rule it states, and the library embeds exactly
the patched file."""
import os
import re
import struct
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "omega3d_b200", "lib")
sys.path.insert(0, os.path.join(ROOT, "tools"))

KERNELS = ["_ZN3o3d10pp2_kernelILi2ELb1ELi384EEEvNS_6PPArgsE", "_ZN3o3d10pp2_kernelILi4ELb0ELi384EEEvNS_6PPArgsE",
           "_ZN3o3d10pp2_kernelILi2ELb1ELi128EEEvNS_6PPArgsE", "_ZN3o3d10pp2_kernelILi4ELb0ELi128EEEvNS_6PPArgsE"]   # + the small-system CTAs
# ... and the alternate-core kernels (csrc/biot_pp_cores.cuh): ppc_kernel<core 1..3, T, grad, 384>
KERNELS += [f"_ZN3o3d10ppc_kernelILi{c}ELi{t}ELb{g}ELi384EEEvNS_6PPArgsE" for c in (1, 2, 3) for t, g in ((2, 1), (4, 0))]


@pytest.fixture(scope="module")
def cubins():
    base, tuned = os.path.join(LIBDIR, "pp2_base.cubin"), os.path.join(LIBDIR, "pp2_tuned.cubin")
    if not (os.path.exists(base) and os.path.exists(tuned)):
        from omega3d_b200 import _lib
        _lib.build()
    return open(base, "rb").read(), open(tuned, "rb").read(), base, tuned


def test_patch_touches_only_reuse_and_yield_bits(cubins):
    import sass_patch as P
    a, b, base, _ = cubins
    assert len(a) == len(b)
    secs = P.elf_text_sections(a)
    for k in KERNELS:
        assert k in secs
    allowed = (1 << 45) | (0x7 << 58)
    text = [(off, off + size) for off, size in secs.values()]
    diff = [i for i in range(len(a)) if a[i] != b[i]]
    assert diff, "the post-pass changed nothing"
    words = sorted({i // 8 * 8 for i in diff})
    changed = 0
    for w in words:
        lo = next((lo for lo, hi in text if lo <= w < hi), None)
        assert lo is not None, "a byte outside the .text sections changed"
        assert (w - lo) % 16 == 8, "a byte outside the upper (control) word of an instruction changed"
        x = struct.unpack_from("<Q", a, w)[0] ^ struct.unpack_from("<Q", b, w)[0]
        assert x & ~allowed == 0
        assert struct.unpack_from("<Q", b, w)[0] & (1 << 45), "reuse flag on an instruction that still carries a yield hint"
        changed += 1
    assert changed >= 20


def test_every_added_flag_has_a_consumer(cubins):
    """In the patched listing: an operand flagged .reuse is read in the same slot by the next instruction (or was flagged
    by ptxas itself), and is never a register the flagged instruction overwrites."""
    import sass_patch as P
    _, _, base, tuned = cubins
    before, after = P.sass(base), P.sass(tuned)
    added = 0
    for k in KERNELS:
        ib, ia = before[k], after[k]
        assert [re.sub(r"\.reuse", "", t) for _, t in ib] == [re.sub(r"\.reuse", "", t) for _, t in ia], "instructions or operands changed"
        for n, ((_, tb), (_, ta)) in enumerate(zip(ib, ia)):
            if tb == ta:
                continue
            o0, o1 = P.operands(ta), P.operands(ia[n + 1][1])
            ob = P.operands(tb)
            assert o0 is not None and o1 is not None
            for slot, (r, wide, flagged) in o0[1].items():
                if flagged and not ob[1][slot][2]:
                    added += 1
                    assert wide and o1[1][slot][0] == r and o1[1][slot][1]
                    assert not (o0[0] <= r <= o0[0] + 1)
    assert added >= 20


def test_library_embeds_the_patched_cubin(cubins):
    _, b, _, _ = cubins
    lib = open(os.path.join(LIBDIR, "libo3d_cuda.so"), "rb").read()
    assert lib.find(b) >= 0, "libo3d_cuda.so does not contain lib/pp2_tuned.cubin byte for byte"


def test_model_counts_fewer_third_reads_after_the_patch(cubins):
    import sass_rf_model as M
    _, _, base, tuned = cubins
    res = []
    for path in (base, tuned):
        ins = M.kernel_sass(path, "pp2_kernelILi2ELb1ELi384")
        j, i = M.hot_loop(ins)
        res.append(M.model([t for _, t in ins[j:i + 1]]))
    assert res[0][0] == res[1][0]            # same packed instructions per trip
    assert res[1][1] < res[0][1]             # fewer three-read instructions
