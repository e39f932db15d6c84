"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol akua_pbf.h declares,
struct layouts match the reference's, and the product has no CPU fallback."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(akua_lib):
    header = (REPO / "include" / "akua_pbf.h").read_text()
    declared = set(re.findall(r"\b(akua_(?:pbf|slab)_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(akua_lib, name), f"libakua_pbf.so does not export {name}"
    from akuaengine_b200 import ABI_SYMBOLS
    assert declared == set(ABI_SYMBOLS)
    assert akua_lib.akua_pbf_abi_version() == 1


def test_struct_layouts_mirror_reference(akua_lib):
    from akuaengine_b200 import PARTICLE_DTYPE, LambdaCorrParams, PBFConfig
    # include/AkuaEngine/Simulation/Particle.h:8-31 offsets
    want = {"position": 0, "velocity": 12, "new_position": 24, "new_velocity": 36, "position_delta": 48, "vorticity": 60,
            "mass": 72, "density": 76, "lambda": 80, "hash": 84, "color": 88, "size": 104}
    assert PARTICLE_DTYPE.itemsize == 108
    for k, off in want.items():
        assert PARTICLE_DTYPE.fields[k][1] == off
    assert C.sizeof(PBFConfig) == 48 and C.sizeof(LambdaCorrParams) == 16
    # defaults come from the library and equal the reference's (PBFConfig.h:10-29)
    cfg, corr = PBFConfig(), LambdaCorrParams()
    c2, k2 = PBFConfig(), LambdaCorrParams()
    akua_lib.akua_pbf_default_config(C.byref(c2))
    akua_lib.akua_pbf_default_corr(C.byref(k2))
    for f, _ in PBFConfig._fields_:
        a, b = getattr(cfg, f), getattr(c2, f)
        assert (list(a) == list(b)) if f == "gravity" else (a == b), f
    assert (c2.restDensity, c2.smoothRadius, c2.maxNeighbours, c2.solverIterations) == (7600.0, np.float32(0.1), 128, 4)
    assert list(c2.gravity) == [0.0, np.float32(-9.8), 0.0]
    assert (k2.enabled, k2.k, k2.n, k2.delta_q) == (1, np.float32(1e-4), 4.0, np.float32(0.03))
    assert (corr.k, corr.n) == (k2.k, k2.n)


def test_invalid_arguments_are_rejected_without_a_device(akua_lib):
    from akuaengine_b200 import LambdaCorrParams, PBFConfig
    h = C.c_void_p()
    cfg, corr = PBFConfig(), LambdaCorrParams()
    assert akua_lib.akua_pbf_create(None, 10, C.byref(cfg), C.byref(corr), None) == 1
    assert akua_lib.akua_pbf_create(C.byref(h), -1, C.byref(cfg), C.byref(corr), None) == 1
    bad = PBFConfig(spatialHashCellSize=0.2)  # reference silently breaks when cell size != smoothing radius
    assert akua_lib.akua_pbf_create(C.byref(h), 10, C.byref(bad), C.byref(corr), None) == 1
    assert akua_lib.akua_pbf_step(None, 0.01, None, None) == 1


def test_no_cpu_fallback():
    """Without a CUDA device construction must fail loudly (never route through oracle/ or any CPU path)."""
    from conftest import HAS_GPU
    from akuaengine_b200 import AkuaError, PBFSolver
    if HAS_GPU:
        pytest.skip("GPU present")
    with pytest.raises(AkuaError):
        PBFSolver(100)


def test_product_never_imports_oracle():
    for f in list((REPO / "akuaengine_b200").rglob("*.py")) + list((REPO / "akuaengine_b200" / "csrc").glob("*")):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "pbf_oracle" not in txt, f


def test_missing_library_is_loud(tmp_path):
    from akuaengine_b200 import AkuaError, load_library
    with pytest.raises(AkuaError):
        load_library(tmp_path / "nope.so")


@pytest.mark.gpu
def test_graphics_resource_export_rejects_a_null_handle(akua_lib):
    """akua_pbf_export_to_graphics_resource needs a registered cudaGraphicsResource (a GL context, not available headless):
    only its argument checking can run here."""
    from akuaengine_b200 import PBFSolver
    s = PBFSolver(64)
    assert akua_lib.akua_pbf_export_to_graphics_resource(s._h, None) == 1
    assert b"null resource" in akua_lib.akua_pbf_last_error(s._h)
    s.close()


@pytest.mark.gpu
def test_device_export_honours_the_renderer_layout(akua_lib):
    """the GL consumer binds position @0, color @88, size @104 at stride 108
    (src/Rendering/Renderer.cpp:201-213). akua_pbf_export_to_graphics_resource maps a registered VBO and runs exactly
    akua_pbf_export_aos108_device on the mapped pointer; without a GL / EGL context on the box the mapped pointer is stood in
    for by a plain device allocation (the export cannot tell the difference: it only sees a device pointer). The exported
    buffer must equal the host download byte for byte and carry the payload at the offsets the renderer reads."""
    import torch
    from akuaengine_b200 import PARTICLE_DTYPE, PBFSolver, scenes
    p, bmin, bmax = scenes.dam_break(20)
    n = len(p)
    rng = np.random.default_rng(3)
    p["color"] = rng.random((n, 4), dtype=np.float32)
    p["size"] = rng.uniform(10, 90, n).astype(np.float32)
    s = PBFSolver(n)
    s.upload_particles(p)
    for _ in range(3):
        s.step(0.0083, bmin, bmax)
    vbo = torch.zeros(n * 108, dtype=torch.uint8, device="cuda")       # the "mapped VBO"
    assert akua_lib.akua_pbf_export_aos108_device(s._h, vbo.data_ptr(), n) == 0
    s.sync()
    host = s.download_particles()
    raw = vbo.cpu().numpy()
    assert raw.tobytes() == host.tobytes()
    got = raw.view(PARTICLE_DTYPE)
    rec = raw.reshape(n, 108)
    assert np.array_equal(rec[:, 0:12].copy().view(np.float32).reshape(n, 3), got["position"])      # attribute 0
    assert np.array_equal(rec[:, 88:104].copy().view(np.float32).reshape(n, 4), got["color"])       # attribute 1
    assert np.array_equal(rec[:, 104:108].copy().view(np.float32).reshape(n), got["size"])          # attribute 2
    # the payload followed its particle through three sorts: match by the upload index carried in a colour channel
    pos4, _, pid = s.download()
    assert np.array_equal(got["color"], p["color"][pid]) and np.array_equal(got["size"], p["size"][pid])
    assert np.array_equal(got["position"], pos4[:, :3])
    # a buffer that is too small or a wrong count is refused, not overrun
    assert akua_lib.akua_pbf_export_aos108_device(s._h, vbo.data_ptr(), n - 1) == 1
    s.close()


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The Python binding re-declares the header's structs by hand; a C program compiled against include/akua_pbf.h prints
    sizeof / offsetof of every field and the ctypes mirrors must agree (catches drift when an option is added)."""
    import subprocess
    from akuaengine_b200 import Counters, LambdaCorrParams, PBFConfig, PBFOptions
    mirrors = {"akua_pbf_config": PBFConfig, "akua_corr_params": LambdaCorrParams, "akua_pbf_options": PBFOptions,
               "akua_pbf_counters": Counters}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "akua_pbf.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    res = subprocess.run(["gcc", "-std=c11", "-I", str(REPO / "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr       # also proves the header is plain C and every mirrored field exists in it
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l}
    for cname, cls in mirrors.items():
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
