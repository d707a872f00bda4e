"""CPU checks of the C-ABI boundary: the shared library loads and exports every
symbol include/ga_b200.h declares; host-only entry points behave; the product
package never imports the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ga_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ga_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gaussiananything_b200 import _lib
    L = _lib.lib()
    names = _declared_functions()
    assert len(names) >= 7
    for n in names:
        assert hasattr(L, n), "libga_b200.so does not export %s" % n
    assert b"sm_100a" in L.ga_b200_version()


def test_layout_is_host_only_and_monotonic():
    from gaussiananything_b200 import _lib
    L = _lib.lib()
    a, b = _lib.GaRasterLayout(), _lib.GaRasterLayout()
    assert L.ga_raster_layout(1, 1000, 2, 64, 64, 5000, C.byref(a)) == 0
    assert L.ga_raster_layout(1, 1000, 2, 64, 64, 10000, C.byref(b)) == 0
    assert b.total_bytes > a.total_bytes > 0
    offs = [a.status, a.rec, a.depth, a.rect, a.tile_count, a.tile_start, a.keys, a.ids, a.final_T, a.n_contrib]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    # error behaviour: bad sizes are rejected, not crashed on
    assert L.ga_raster_layout(0, 10, 1, 64, 64, 10, C.byref(a)) == -1
    assert L.ga_raster_layout(1, 10, 1, 5000, 64, 10, C.byref(a)) == -3
    assert L.ga_raster_backward_scratch_bytes(1, 1000, 2) >= 1000 * 2 * 18 * 4


def test_forward_rejects_null_and_small_workspace():
    from gaussiananything_b200 import _lib
    L = _lib.lib()
    assert L.ga_raster_forward(None, 1, 10, 1, None, None, None, 32, 32, 1.0, None, None, None, None, 0, 10, None) == -1


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "gaussiananything_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "from oracle" not in s, f
                assert "surfel_oracle.so" not in s and "libsurfel_oracle" not in s, f


def test_cpu_tensors_fail_loudly():
    import torch
    from gaussiananything_b200 import raster
    with pytest.raises(RuntimeError):
        raster.forward_raw(torch.zeros(1, 4, 13), torch.eye(4).reshape(1, 1, 4, 4),
                           torch.eye(4).reshape(1, 1, 4, 4), torch.ones(3), 32, 32)


def test_install_shims_registers_reference_module_names():
    import sys
    import gaussiananything_b200 as ga
    saved = {k: sys.modules.get(k) for k in ("diff_surfel_rasterization", "transport", "transport.transport")}
    try:
        ga.install_shims()
        import diff_surfel_rasterization as dsr          # the name nsr/gs_surfel.py:15 imports
        import transport as tr                           # the name flow_matching_trainer.py imports
        assert hasattr(dsr, "GaussianRasterizationSettings") and hasattr(dsr, "GaussianRasterizer")
        assert hasattr(tr, "create_transport") and hasattr(tr, "Sampler")
        fields = dsr.GaussianRasterizationSettings._fields
        assert fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                          "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_missing_library_fails_loudly_everywhere():
    """No CPU fallback anywhere on the product path: with the shared library absent the rasteriser, the DiT binding and
    the VAE decoder all raise (checked in a child process so this process keeps its loaded library)."""
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from gaussiananything_b200 import _lib, raster, dit, vae_decoder\n"
        "n = 0\n"
        "for f in (_lib.lib, dit._bind, vae_decoder._bind,\n"
        "          lambda: raster.layout(1, 10, 1, 32, 32, 100)):\n"
        "    try:\n"
        "        f()\n"
        "    except RuntimeError as e:\n"
        "        assert 'no CPU fallback' in str(e) or 'missing' in str(e), e\n"
        "        n += 1\n"
        "print('raised', n)\n")
    env = dict(os.environ, GA_B200_LIB="/nonexistent/libga_b200.so")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "raised 4" in out.stdout, out.stdout + out.stderr[-1000:]


def test_decoder_refuses_cpu_devices():
    """(The DiT modules' CPU refusal is in tests/test_oracle_dit.py::test_dit_module_has_reference_state_dict_layout.)"""
    from gaussiananything_b200.vae_decoder import SurfelDecoder
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SurfelDecoder({}, 12, 12, device="cpu")
