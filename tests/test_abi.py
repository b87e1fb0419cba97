"""The C-ABI shared library loads without a GPU and exports every symbol the header
declares; without a device every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest


def _gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol(ffi):
    L = ffi.lib()
    declared = ffi.header_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert set(declared) == set(ffi._SIGS), set(declared) ^ set(ffi._SIGS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", ffi.LIB_PATH]).decode()
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared) <= exported


def test_stats_struct_layout_matches_the_header(ffi, tmp_path):
    # the ctypes mirror of vidx_search_stats must have the C compiler's layout: field offsets and size
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields = [n for n, _ in ffi.SearchStats._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vidx_b200.h"\nint main(void) {\n'
                   + "".join(f'  printf("{n} %zu\\n", offsetof(vidx_search_stats, {n}));\n' for n in fields)
                   + '  printf("sizeof %zu\\n", sizeof(vidx_search_stats));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    out = dict(ln.split() for ln in subprocess.check_output([str(exe)]).decode().splitlines())
    assert int(out.pop("sizeof")) == C.sizeof(ffi.SearchStats)
    assert {n: int(v) for n, v in out.items()} == {n: getattr(ffi.SearchStats, n).offset for n in fields}


def test_library_is_sm100a_cuda_code(ffi):
    out = subprocess.run(["cuobjdump", "-lelf", ffi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:500]


def test_product_does_not_reference_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "vector-indexer_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".py", "Makefile")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "vidx_oracle" not in text and "import oracle" not in text and "oracle/" not in text.replace(
                    "test oracle under oracle/", ""), os.path.join(dp, f)


def test_host_only_entry_points(ffi):
    assert ffi.calculate_num_clusters(50000) == 448 and ffi.calculate_num_clusters(10 ** 6) == 4000
    assert ffi.calculate_max_iterations(5000) == 300 and ffi.calculate_max_iterations(10 ** 6) == 20
    ix = ffi.Index(16)
    assert ix.dimension == 16 and ix.ntotal == 0 and ix.nlist == 0
    ix.set_limits(10, 20, 100, 100)
    with pytest.raises(ffi.VidxError):
        h = C.c_void_p()
        ffi.check(ffi.lib().vidx_create(0, 0, C.byref(h)))  # dimension 0


def test_invalid_input_is_reported_before_any_device_work(ffi):
    ix = ffi.Index(8)
    with pytest.raises(ffi.InvalidInput):
        ix.build(np.zeros((0, 8), np.float32))          # "no vectors provided" (api.rs:116-118)
    with pytest.raises(ffi.InvalidInput):
        ix.search(np.zeros((2, 8), np.float32), 0, 4)   # ivf_index.rs:197-202
    with pytest.raises(ffi.InvalidInput):
        ix.search(np.zeros((2, 8), np.float32), 4, 0)
    with pytest.raises(ffi.InvalidInput):
        ffi.kmeans_mini_batch(np.zeros((0, 8), np.float32), 3, 5)  # kmeans.rs:72-77


@pytest.mark.skipif(_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_a_gpu(ffi):
    ix = ffi.Index(8)
    x = np.random.default_rng(0).standard_normal((100, 8)).astype(np.float32)
    with pytest.raises(ffi.VidxError) as e:
        ix.build(x)
    assert e.value.code == ffi.CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(ffi.VidxError) as e:
        ffi.kmeans_mini_batch(x, 4, 5)
    assert e.value.code == ffi.CUDA
    with pytest.raises(ffi.VidxError):  # nothing built -> error, never a silent empty answer
        ix.search(x[:2], 3, 2)


def test_python_package_surface():
    # bindings/python/python/vector_indexer_py/__init__.py:23
    import vector_indexer_py as vip
    assert set(vip.__all__) == {"build", "load", "suggest_nlist", "VectorIndex"}
    assert vip.suggest_nlist(9999) == 99 and vip.suggest_nlist(10000) == 200
    with pytest.raises(RuntimeError):
        vip.build(np.zeros((0, 4), np.float32))
    with pytest.raises(RuntimeError):
        vip.load("/nonexistent/index", "/nonexistent/shards", 4)
