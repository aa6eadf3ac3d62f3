"""CPU: the C-ABI library builds, loads, and exports every symbol include/ddrl_b200.h declares; the ctypes
prototypes cover the header one to one; pure-host entry points behave (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ddrl_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddrl_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = header_symbols()
    for must in ("ddrl_fcnet_forward", "ddrl_ppo_train_step", "ddrl_grad_reduce", "ddrl_clip_adam", "ddrl_filter_update",
                 "ddrl_filter_partial", "ddrl_filter_merge", "ddrl_gae", "ddrl_adv_standardize", "ddrl_gather_rows",
                 "ddrl_ppo_loss_grad", "ddrl_graphnet_forward", "ddrl_graphnet_backward", "ddrl_gcn_forward",
                 "ddrl_leg_coupling", "ddrl_dg_sample", "ddrl_last_error", "ddrl_launch_count", "ddrl_fcnet_pack",
                 "ddrl_fcnet_image_floats", "ddrl_umma_selftest"):
        assert must in syms


def test_library_exports_every_header_symbol():
    from ddrl_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in ddrl_b200.h but not exported by libddrl_b200.so"


def test_ctypes_prototypes_cover_header_exactly():
    from ddrl_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == header_symbols()


def test_no_torch_types_in_the_abi():
    raw = open(HEADER).read()
    assert 'extern "C"' in raw
    src = re.sub(r"/\*.*?\*/", "", raw, flags=re.S)            # declarations only; comments may mention torch
    assert "torch" not in src.lower() and "at::" not in src and "std::" not in src and "Tensor" not in src


def test_pure_host_entry_points():
    from ddrl_b200 import _lib
    lib = _lib.load()
    assert lib.ddrl_abi_version() == 1
    # parameter-count closed form 128*D + 130*A + 8513 for every published architecture
    for D, A, n in ((43, 8, 15057), (19, 2, 11205), (35, 2, 13253), (27, 2, 12229), (27, 4, 12489), (44, 8, 15185),
                    (20, 2, 11333), (36, 2, 13381), (28, 4, 12617)):
        assert lib.ddrl_fcnet_num_params(D, A) == n == 128 * D + 130 * A + 8513
    assert lib.ddrl_fcnet_num_params(65, 2) == -2 and lib.ddrl_fcnet_num_params(19, 9) == -2   # DDRL_E_UNSUPPORTED_SHAPE
    assert lib.ddrl_graphnet_num_params(4) == 14532 + 14337
    assert lib.ddrl_graphnet_num_params(3) == -2
    assert lib.ddrl_filter_num_partials(1) == 1 and lib.ddrl_filter_num_partials(131072) == 256
    assert lib.ddrl_filter_ws_bytes(4, 131072, 19) == 4 * 256 * 19 * 3 * 8
    assert lib.ddrl_gae_ws_bytes(4, 4096) == 4 * 16 * 2 * 8
    assert isinstance(lib.ddrl_launch_count(), int)


def test_bad_arguments_return_error_codes_without_touching_a_gpu():
    from ddrl_b200 import _lib
    lib = _lib.load()
    assert lib.ddrl_fcnet_forward(None, None, None, None, 0.0, 1, 1, 19, 2, None, None, None, None, None, None, None) == -1
    assert b"fcnet_forward" in lib.ddrl_last_error()
    assert lib.ddrl_gae(None, None, None, None, 1, 1, 1, 1, 0.99, 0.95, None, None, None, None, None) == -1
    assert lib.ddrl_clip_adam(None, None, None, None, None, 1, 1, 0.0, 0.0, 0.0, 0.0, 0.0, None, None, None, None, None, 0, 0, None) == -1


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ddrl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ddrl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.DDRLError, match="no CPU"):
        _lib.load()


def test_ops_refuse_cpu_tensors():
    import torch
    from ddrl_b200 import kernels as K
    from ddrl_b200._lib import DDRLError
    with pytest.raises(DDRLError, match="CUDA"):
        K.gather_rows(torch.zeros(1, 4, 2), torch.zeros(1, 4, dtype=torch.int32))


def test_plain_c_program_links_and_calls_the_library(tmp_path):
    """The drop-in boundary is a C ABI: a C99 program including only include/ddrl_b200.h links against libddrl_b200.so and
    calls it (sizes, parameter counts, argument validation with ddrl_last_error) — no Python, torch or CUDA headers."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "ddrl_b200")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(lib_dir, "libddrl_b200.so")):
        pytest.skip("gcc or the built library is missing")
    exe = str(tmp_path / "c_abi_smoke")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                         os.path.join(root, "tests", "c_abi_smoke.c"), "-L", lib_dir, "-lddrl_b200", "-o", exe],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, LD_LIBRARY_PATH=lib_dir))
    assert run.returncode == 0, run.stderr
    lines = run.stdout.splitlines()
    assert lines[1] == "fcnet 11205 15057 -2"            # 128 D + 130 A + 8513; D = 65 > DDRL_MAX_OBS -> DDRL_E_UNSUPPORTED_SHAPE
    assert lines[2] == "graphnet 28869"                  # actor 14 532 + critic 14 337
    assert lines[5].startswith("badarg -1 fcnet_forward:")
    assert lines[6] == "sizes 20 8"
    # the ctypes mirror of ddrl_sgd_tail (ddrl_b200/_lib.py) has the layout the C compiler gives the header's struct
    from ddrl_b200._lib import SgdTail
    assert lines[7] == "tail %d %d %d %d %d" % (ctypes.sizeof(SgdTail), SgdTail.lr.offset, SgdTail.peer_x.offset,
                                                 SgdTail.ll_ws.offset, SgdTail.grad_acc.offset)
