"""CPU-only: the C-ABI shared library loads, exports every symbol include/meso_cuda.h declares, agrees with the header on
struct sizes, and fails loudly (no CPU fallback) when there is no GPU.  No compute is called here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "meso_cuda.h")


def _declared():
    src = open(HEADER).read()
    return re.findall(r"MESO_API\s+[\w\s\*]+?\b(meso_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    from mesoengine_b200 import capi
    names = _declared()
    assert len(names) >= 30 and len(set(names)) == len(names)
    assert sorted(names) == sorted(capi.SYMBOLS), set(names) ^ set(capi.SYMBOLS)
    lib = C.CDLL(capi.SO_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.meso_abi_version() == 1


def test_struct_sizes_match_reference_layouts(tmp_path):
    """Compile the header with plain gcc (it must be valid C) and compare sizeof() with the numpy mirrors."""
    from mesoengine_b200 import capi
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include "meso_cuda.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(MesoGPUBlock),sizeof(MesoGPUChunk),sizeof(MesoGPUUniformCamera),sizeof(MesoGPUUniformSceneConfig),'
                    'sizeof(MesoHitRecord),sizeof(MesoQuad),sizeof(MesoRaySetup),sizeof(MesoRayStats));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [12, 16, 160, 16, 16, 16, 80, 120]  # FGPUBlock 12 B, FGPUChunk 16 B, FGPUUniformCamera 160 B, scene 16 B
    assert [capi.GPUBlock.itemsize, capi.GPUChunk.itemsize, capi.Camera.itemsize, capi.SceneConfig.itemsize,
            capi.HitRecord.itemsize, capi.Quad.itemsize, capi.RaySetup.itemsize, capi.RayStats.itemsize] == sizes


def test_pure_host_entry_points_work_without_gpu():
    from mesoengine_b200 import capi
    assert capi.tiles_per_rank(3840, 2160, 8) == 4050
    assert capi.tiles_per_rank(101, 37, 3) == (4 * 5 + 2) // 3
    cam = np.zeros(1, dtype=capi.Camera)
    with pytest.raises(capi.MesoError):
        capi.ray_setup(cam, (0, 0, 0), 0, 10)   # ArgumentOutOfRange-class error, message through meso_last_error()


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run; with one, this test has nothing to check."""
    import torch
    from mesoengine_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MesoError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load oracle/."""
    pkg = os.path.join(ROOT, "mesoengine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libmeso_oracle" not in src and "import orc" not in src and "meso_oracle.h" not in src, os.path.join(dirpath, f)
    out = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(pkg, "libmeso_b200.so")], capture_output=True, text=True).stdout
    assert "orc_" not in out
