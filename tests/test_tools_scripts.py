"""The GPU-box scripts under tools/ cost GPU minutes when they are wrong: every one must parse, and the round-2 scripts (tools/r2_*.sh) may only call
binaries that tools/Makefile / apps/Makefile build, python files that exist, and test files that exist."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SCRIPTS = sorted((ROOT / "tools").glob("*.sh"))


@pytest.mark.parametrize("script", SCRIPTS, ids=lambda p: p.name)
def test_script_parses(script):
    r = subprocess.run(["bash", "-n", str(script)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.parametrize("script", [s for s in SCRIPTS if s.name.startswith("r2_")], ids=lambda p: p.name)
def test_round2_scripts_reference_things_that_exist(script):
    text = script.read_text()
    tools_mk = (ROOT / "tools" / "Makefile").read_text()
    apps_mk = (ROOT / "apps" / "Makefile").read_text()
    for name in set(re.findall(r"\./build/([A-Za-z0-9_\-]+)", text)):
        assert f"$(OUT)/{name}" in tools_mk, f"{script.name}: build/{name} is not a target of tools/Makefile"
    for name in set(re.findall(r"(?<![\w/])bin/([A-Za-z0-9_\-]+)", text)):
        assert name in apps_mk, f"{script.name}: bin/{name} is not built by apps/Makefile"
    for path in set(re.findall(r"python (?:-m pytest )?((?:tools|tests)/[A-Za-z0-9_/]+\.py)", text)) | set(re.findall(r"(tests/[A-Za-z0-9_]+\.py)", text)):
        assert (ROOT / path).exists(), f"{script.name}: {path} does not exist"
    for sh in set(re.findall(r"bash (tools/[A-Za-z0-9_]+\.sh)", text)):
        assert (ROOT / sh).exists(), f"{script.name}: {sh} does not exist"
    for key in set(re.findall(r'-k "([^"]+)"', text)):   # pytest -k expressions must select something
        if "$" in key:
            continue
        r = subprocess.run(["python", "-m", "pytest", "tests/test_experimental_gpu.py", "tests/test_gemm_gpu.py", "--collect-only", "-q", "-k", key], capture_output=True,
                           text=True, cwd=ROOT, timeout=600)
        assert re.search(r"(\d+)/\d+ tests collected", r.stdout) and int(re.search(r"(\d+)/\d+ tests collected", r.stdout).group(1)) > 0, (script.name, key, r.stdout[-300:])


def test_every_switch_the_round2_scripts_set_is_read_somewhere():
    """A misspelt TMM_* variable silently measures the default twice."""
    sources = ""
    for pattern in ("tiled-mm_b200/csrc/*", "tiled-mm_b200/*.py", "tiled_mm_b200.py", "tools/*.py", "tools/*.cu", "tests/*.py", "apps/*"):
        for f in ROOT.glob(pattern):
            if f.is_file():
                sources += f.read_text(errors="ignore")
    missing = []
    for script in SCRIPTS:
        if not script.name.startswith("r2_"):
            continue
        for name in set(re.findall(r"\b(TMM_[A-Z0-9_]+)=", script.read_text())):
            if f'"{name}"' not in sources and f"'{name}'" not in sources:
                missing.append((script.name, name))
    assert not missing, missing
