"""The C-ABI driven by a C compiler: tests/abi_smoke.c includes include/qrochet_b200.h, is compiled by build.sh with
`gcc -std=c99 -Wall -Werror` and walks create -> tensors -> contract -> MPS -> canonize! -> evolve! -> overlap -> expect
(the header proven by C, not by ctypes prototypes)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "abi_smoke")


@pytest.mark.gpu
def test_c_program_drives_the_header_on_the_gpu():
    assert os.path.exists(EXE), "tests/abi_smoke missing: run ./build.sh"
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ABI_SMOKE_OK" in out.stdout


def test_c_program_links_every_symbol_it_uses():
    """CPU half: the executable built from the header loads and resolves its symbols (no compute without a GPU)."""
    assert os.path.exists(EXE), "tests/abi_smoke missing: run ./build.sh"
    out = subprocess.run([EXE, "--symbols"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ABI_SYMBOLS_OK" in out.stdout
