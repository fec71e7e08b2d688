"""CPU: the C-ABI library loads, exports every symbol include/grok_b200.h declares, the ctypes mirror of the
structs has the C layout, and the product fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import grokimagecompression_b200 as gb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "grok_b200.h")).read()
    return sorted(set(re.findall(r"GB200_API[^;(]*?\b(gb200_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = gb.lib()
    names = _declared()
    assert len(names) >= 36
    for n in names:
        assert hasattr(L, n), n
    assert sorted(gb.SYMBOLS) == names
    out = subprocess.check_output(["nm", "-D", "--defined-only", gb.LIB_PATH], text=True)
    exported = set(re.findall(r" T (gb200_\w+)", out))
    assert exported == set(names)
    assert L.gb200_abi_version() == gb.ABI_VERSION == 3


def test_struct_layouts_match_c(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "grok_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(gb200_comp_params),offsetof(gb200_comp_params,stepsize),offsetof(gb200_comp_params,rd_weight),'
                    'sizeof(gb200_tile_params),sizeof(gb200_cblk_info),sizeof(gb200_cblk_enc),sizeof(gb200_cblk_dec),sizeof(gb200_t1_block));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(gb.CompParams), gb.CompParams.stepsize.offset, gb.CompParams.rd_weight.offset, C.sizeof(gb.TileParams),
            gb.CBLK_INFO_DTYPE.itemsize, gb.CBLK_ENC_DTYPE.itemsize, gb.CBLK_DEC_DTYPE.itemsize, gb.T1_BLOCK_DTYPE.itemsize]
    assert got == want


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gb.GrokB200Error) as e:
        gb.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """only tests/, __graft_entry__.smoke() and the CPU-baseline legs of bench.py may touch oracle/: the package, the
    reference-side bindings under integration/ and everything else in bench.py must not"""
    bad = ("gb_oracle", "libgrkref", "gbo_", "oracle_pipeline")
    for top in ("grokimagecompression_b200", "integration", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, f), errors="replace").read()
                    assert not any(b in txt for b in bad), (top, f)
                    if top == "grokimagecompression_b200":
                        assert "oracle/" not in txt, f
    # bench.py: every function that names the oracle tree or its loaders belongs to the reference / cpu_baseline legs
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    allowed = {"cpu_reference_run", "run_reference"}

    def touches(node):
        for n in ast.walk(node):
            if isinstance(n, ast.Import) and any(a.name.split(".")[0] in ("_libs", "oracle_pipeline") for a in n.names):
                return True
            if isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] in ("_libs", "oracle_pipeline"):
                return True
            if isinstance(n, ast.Name) and n.id in ("_libs", "oracle_pipeline"):
                return True
            if isinstance(n, ast.Constant) and isinstance(n.value, str) and ("libgb_oracle" in n.value or "libgrkref" in n.value):
                return True
        return False

    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            assert not touches(node) or node.name in allowed, node.name
        else:
            assert not touches(node), ast.get_source_segment(src, node)[:80]


PLUGIN_SO = os.path.join(ROOT, "integration", "_build", "libgrok_plugin.so")
# what the host resolves by name with dlsym: grok.cpp:810-822, plugin_bridge.cpp:302-303, minpf_plugin_manager.cpp:146-147
PLUGIN_ABI = ["minpf_post_load_plugin", "plugin_init", "plugin_encode", "plugin_batch_encode", "plugin_is_batch_complete",
              "plugin_stop_batch_encode", "plugin_decode", "plugin_init_batch_decode", "plugin_batch_decode", "plugin_stop_batch_decode",
              "plugin_get_debug_state", "plugin_debug_mqc_next_cxd", "plugin_debug_mqc_next_plane"]


@pytest.mark.skipif(not os.path.exists(PLUGIN_SO), reason="integration/_build not built")
def test_plugin_adapter_exports_the_minpf_abi_and_host_falls_back_without_gpu(tmp_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", PLUGIN_SO], text=True)
    exported = set(re.findall(r" T (\w+)", out))
    for n in PLUGIN_ABI:
        assert n in exported, n
    import torch
    if torch.cuda.is_available():
        return
    # the reference's own loader finds, loads and registers the plugin; plugin_init says no (no device), so the host keeps
    # its CPU path: that is status -1 of the driver, never a silently CPU-computed "plugin" result
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, _libs\n"
            "_libs.write_pnm(%r, [np.arange(48 * 64, dtype=np.int32).reshape(48, 64) %% 251], 8)\n"
            "r = _libs.ref_plugin_encode_file(%r, 48 * 64)\n"
            "assert r == -1, r\n") % (os.path.join(ROOT, "tests"), ROOT, str(tmp_path / "a.pgm"), str(tmp_path / "a.pgm"))
    subprocess.check_call([os.sys.executable, "-c", code], timeout=120)
