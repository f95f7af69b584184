"""CPU-side checks of the C-ABI boundary: the shared library loads, exports every symbol that
include/stgcma.h declares, and the product path fails loudly (no fallback) without it / without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "stgcma.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stgcma_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import stgcma.lib as lib
    names = declared_symbols()
    assert len(names) >= 19
    assert os.path.isfile(lib.LIB_PATH), "libstgcma.so must be built in-tree (python -c 'import __graft_entry__ as g; g.build()')"
    h = ctypes.CDLL(lib.LIB_PATH)
    missing = [n for n in names if not hasattr(h, n)]
    assert not missing, missing
    assert sorted(lib.EXPORTS) == names, (sorted(set(lib.EXPORTS) ^ set(names)))
    h.stgcma_version.restype = ctypes.c_int
    assert h.stgcma_version() >= 100


def declared_prototypes():
    """{name: [ctypes type per parameter]} parsed from the C prototypes of include/stgcma.h."""
    src = open(os.path.join(ROOT, "include", "stgcma.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|const char\*)\s+(stgcma_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        types = []
        for a in (x.strip() for x in args.split(",")):
            if a in ("void", ""):
                continue
            if "*" in a:
                types.append(ctypes.c_void_p)
            elif a.startswith("int64_t"):
                types.append(ctypes.c_int64)
            elif a.startswith("float"):
                types.append(ctypes.c_float)
            elif a.startswith("int"):
                types.append(ctypes.c_int)
            else:
                raise AssertionError(f"unparsed parameter {a!r} of {name}")
        protos[name] = types
    return protos


def test_ctypes_signatures_match_the_header():
    """stgcma.lib._SIGNATURES (the argtypes every call goes through) against the
    prototypes of include/stgcma.h, parameter by parameter."""
    import stgcma.lib as lib
    protos = declared_prototypes()
    assert set(lib._SIGNATURES) <= set(protos)
    for name, argtypes in lib._SIGNATURES.items():
        want = protos[name]
        got = list(argtypes)
        assert len(got) == len(want), (name, len(got), len(want))
        for i, (g, w) in enumerate(zip(got, want)):
            assert g is w, f"{name}: parameter {i} is {g.__name__} in lib.py, {w.__name__} in stgcma.h"


def test_pack_entry_matches_struct_layout():
    import stgcma.lib as lib
    assert lib.PACK_ENTRY.size == 48  # sizeof(StgcmaPackEntry): 2 pointers, 2 int32, 2 int64, 2 int32


def test_adam_entry_matches_struct_layout():
    import stgcma.lib as lib
    assert lib.ADAM_ENTRY.size == 48  # sizeof(StgcmaAdamEntry): 4 pointers, int64, 2 int32


def test_fused_adam_refuses_cpu_parameters():
    from stgcma.optim import FusedAdam
    import stgcma.lib as lib
    p = torch.nn.Parameter(torch.zeros(4))
    opt = FusedAdam([p], lr=1e-3)
    p.grad = torch.ones(4)
    with pytest.raises(lib.StgcmaError):
        opt.step()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import stgcma.lib as lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", str(tmp_path / "libstgcma.so"))
    with pytest.raises(lib.StgcmaError, match="no CPU / eager fallback"):
        lib.load()
    with pytest.raises(lib.StgcmaError):
        lib.layernorm_fwd(torch.zeros(4, 8), torch.ones(8), torch.zeros(8), torch.zeros(4, 8), None, None, 4, 8)


def test_cpu_tensors_are_refused():
    import model
    m = model.Swin_AVE.SwinTransformer2D_Adapter_New(label_dim=7, img_size=112, num_frames=2, embed_dim=32, depths=[2, 2],
                                                     num_heads=[1, 2], adapter_mlp_ratio=[0.25, 0.25], ftmode="fusion")
    import stgcma.lib as lib
    with pytest.raises(lib.StgcmaError, match="no CPU fallback"):
        m(torch.zeros(1, 2, 112, 112), torch.zeros(1, 3, 2, 112, 112), "fusion")


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "stg-cma_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)
