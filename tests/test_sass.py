"""What the built library's hot kernels are made of, read from its SASS (cuobjdump, no GPU needed): the FP64 kernels issue DMMA fed by TMA,
the float kernel issues tcgen05 MMAs (UTCHMMA) fed by TMA with TMEM loads, nothing spills to local memory; the opt-in experimental kernels
carry the instructions their design rests on (A operand from tensor memory, CTA-pair MMAs, integer MMAs).  Mnemonics as listed in the
profiling recipe (tcgen05.mma -> UTC*MMA, cp.async.bulk.tensor -> UTMALDG)."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "tiled-mm_b200" / "libtiledmm_b200.so"


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists() or not LIB.exists():
        pytest.skip("cuobjdump or the built library is not available")
    text = subprocess.run([exe, "-sass", str(LIB)], capture_output=True, text=True, timeout=600, check=True).stdout
    funcs, name = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name:
            mm = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if mm:
                funcs[name].append(mm.group(1))
    return funcs


def kernels(sass, *needles):
    out = {k: v for k, v in sass.items() if all(n in k for n in needles)}
    assert out, needles
    return out


def has(ops, prefix):
    return any(o.startswith(prefix) for o in ops)


def no_local_memory(ops):
    return not any(o.startswith(("STL", "LDL")) for o in ops)


def test_fp64_kernels_are_dmma_fed_by_tma(sass):
    for name, ops in {**kernels(sass, "3f64", "dgemm_kernel"), **kernels(sass, "3c64", "zgemm_kernel")}.items():
        assert has(ops, "DMMA"), name
        assert has(ops, "UTMALDG"), name
        assert has(ops, "SYNCS"), name       # mbarrier pipeline
        assert no_local_memory(ops), name    # the 232-register math warps must not spill


def test_float_kernel_is_tcgen05_with_tmem(sass):
    for name, ops in kernels(sass, "f32tc", "sgemm_tc_kernel").items():
        assert has(ops, "UTCHMMA") and has(ops, "UTMALDG") and has(ops, "LDTM") and has(ops, "UTCBAR"), name
        assert not has(ops, "UTCHMMA.2CTA"), name
        assert no_local_memory(ops), name
        assert ops.count("UTCHMMA") >= 12, name   # 4 k-steps x 3 terms per stage


def test_int8_emulation_kernel_carries_its_instructions(sass):
    """FP64 emulation on the integer tensor cores (opt-in TMM_F64_MATH=i8): integer MMAs, int32 -> FP64 conversion and FP64 accumulation."""
    found = kernels(sass, "f64i8", "dgemm_i8_kernel")
    assert found
    for name, ops in found.items():
        assert has(ops, "UTCIMMA") and has(ops, "I2F.F64") and has(ops, "DFMA") and no_local_memory(ops), name
        assert not has(ops, "UTCIMMA.2CTA"), name


def test_hot_kernels_fit_their_occupancy_budget():
    """cuobjdump -res-usage: the TMA / tensor-core kernels are compiled for 128 registers (2 x 256-thread CTAs per SM for the FP64 kernels,
    one 512-thread CTA for the tcgen05 kernels - the warpgroups then re-balance with setmaxnreg) and use no stack and no local memory."""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists() or not LIB.exists():
        pytest.skip("cuobjdump or the built library is not available")
    text = subprocess.run([exe, "-res-usage", str(LIB)], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = text.splitlines()
    seen = 0
    for i, line in enumerate(lines):
        m = re.search(r"Function (\S+):", line)
        if not m or not any(k in m.group(1) for k in ("dgemm_kernel", "zgemm_kernel", "sgemm_tc", "dgemm_i8_kernel")):
            continue
        usage = lines[i + 1]
        reg, stack, local = (int(re.search(rf"{k}:(\d+)", usage).group(1)) for k in ("REG", "STACK", "LOCAL"))
        assert reg <= 128 and stack == 0 and local == 0, (m.group(1), usage)
        seen += 1
    assert seen >= 10
