"""The C-ABI library loads and exports every symbol include/i3d_c_api.h declares (no compute calls)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "i3d_c_api.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(i3d_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from intrinsic3d_b200 import engine
    L = engine.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in i3d_c_api.h but not exported"
    assert sorted(engine.EXPORTED_SYMBOLS) == names


def test_struct_sizes_and_defaults():
    from intrinsic3d_b200 import engine
    from intrinsic3d_b200.ctypes_defs import I3DIterInfo, I3DParams, default_params
    L = engine.load_library()
    assert L.i3d_abi_version() == 1
    assert L.i3d_sizeof_params() == C.sizeof(I3DParams)
    assert L.i3d_sizeof_iter_info() == C.sizeof(I3DIterInfo)
    a, b = engine.default_params(), default_params()
    assert bytes(a) == bytes(b)
    assert a.num_observations == 5 and a.lm_steps == 50 and a.max_linear_solver_iterations == 500
    from intrinsic3d_b200.ctypes_defs import I3DLightingInfo, I3DLightingParams
    import oracle
    L.i3d_sizeof_lighting_params.restype = C.c_uint64
    L.i3d_sizeof_lighting_info.restype = C.c_uint64
    assert L.i3d_sizeof_lighting_params() == C.sizeof(I3DLightingParams)
    assert L.i3d_sizeof_lighting_info() == C.sizeof(I3DLightingInfo)
    la, lb = engine.default_lighting_params(), oracle.default_lighting_params()
    assert bytes(la) == bytes(lb)
    assert abs(la.subvolume_size - 0.2) < 1e-7 and la.lambda_reg == 10.0 and la.max_iterations == 50


def test_no_cpu_fallback():
    """Without a GPU the engine must refuse loudly; with one, creation works."""
    import torch
    from intrinsic3d_b200 import engine
    L = engine.load_library()
    h = C.c_void_p()
    rc = L.i3d_engine_create(C.c_int(0), C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        L.i3d_engine_destroy(h)
    else:
        assert rc != 0 and not h.value
        msg = L.i3d_last_error(None).decode()
        assert "no CPU fallback" in msg or "CUDA" in msg


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "intrinsic3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
