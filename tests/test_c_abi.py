"""The C ABI from plain C: include/hakai_b200.h compiles as pedantic C99 and a C host links against
libhakai_b200.so and drives it without Python.  In the GPU-less build container hk_create must refuse (there is no
CPU path); the same binary runs a one-element model on a GPU box (tests/c_abi/abi_driver.c)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "hakai_fem_b200")


def _build(tmp_path):
    exe = str(tmp_path / "abi_driver")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror",
                           os.path.join(ROOT, "tests", "c_abi", "abi_driver.c"), "-I", os.path.join(ROOT, "include"),
                           "-L", LIBDIR, "-lhakai_b200", "-lm", "-Wl,-rpath," + LIBDIR, "-o", exe])
    return exe


@pytest.mark.skipif(not os.path.exists(os.path.join(LIBDIR, "libhakai_b200.so")), reason="library not built")
def test_header_is_plain_c_and_library_links_and_refuses_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked run")
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.startswith("NO_DEVICE"), out.stdout


def test_c_driver_logic_on_the_host_compiled_kernels(tmp_path):
    """The same C program against tests/emu/libhakai_emu.so (macro-renamed entry points): its one-element run and its
    own assertions hold, so the GPU run of the unmodified binary checks the CUDA library, not the driver."""
    from .emu import emu_engine
    emu_engine.build()
    emu_dir = os.path.dirname(emu_engine.LIB_PATH)
    exe = str(tmp_path / "abi_driver_emu")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-include",
                           os.path.join(ROOT, "tests", "c_abi", "emu_rename.h"),
                           os.path.join(ROOT, "tests", "c_abi", "abi_driver.c"), "-I", os.path.join(ROOT, "include"),
                           "-L", emu_dir, "-lhakai_emu", "-lm", "-Wl,-rpath," + emu_dir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr + out.stdout
    assert out.stdout.startswith("OK u_z=2.000000e-04"), out.stdout
