"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the symbols that
include/nerf_b200.h declares; argument errors are reported without a GPU; the host mirror keeps the reference's
names.  No compute is launched here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "nerf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nerf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nerf_b200.h but not exported"
    assert set(pkg._lib.SIGNATURES) == set(names), "ctypes table and header disagree"


def test_version_and_geometry(pkg):
    lib = pkg.load()
    assert b"sm_100a" in lib.nerf_version()
    cfg = pkg.NetCfg(5, 4, 2, 256, 128, 0.05)
    assert lib.nerf_param_count(ctypes.byref(cfg)) == 514332          # 514 332 params/net
    assert lib.nerf_xyz_enc_dim(ctypes.byref(cfg)) == 33
    assert lib.nerf_view_enc_dim(ctypes.byref(cfg)) == 24
    cfg0 = pkg.NetCfg(5, 4, 0, 256, 128, 0.05)
    assert lib.nerf_view_enc_dim(ctypes.byref(cfg0)) == 0
    from oracle import nerf_oracle as O
    assert lib.nerf_param_count(ctypes.byref(cfg0)) == O.NetCfg(5, 4, 0).n_params
    cfg1 = pkg.NetCfg(5, 2, 1, 256, 128, 0.05)
    assert lib.nerf_view_enc_dim(ctypes.byref(cfg1)) == 8


def test_argument_errors_are_reported(pkg):
    lib = pkg.load()
    # n_angles_for_model outside {1,2}: the reference raises Exception('... should be 1 or 2.') (src/UtilsCV.py:138)
    st = lib.nerf_view_directions(ctypes.c_void_p(16), 1, 1, 3, ctypes.c_void_p(16), None)
    assert st == -1 and b"should be 1 or 2" in lib.nerf_last_error()
    st = lib.nerf_composite_fwd(None, None, 1, 64, None, None, None, None, None, None, None, None)
    assert st == -1
    bad = pkg.NetCfg(5, 4, 7, 256, 128, 0.05)
    assert lib.nerf_param_count(ctypes.byref(bad)) < 0


def test_host_mirror_keeps_reference_names(pkg):
    ucv, unrf = pkg.UtilsCV, pkg.UtilsNeuralRadianceField
    for name in ("get_rays_directions", "get_z_values", "get_z_vals_from_prob_dist_func", "sample_along_rays",
                 "get_view_directions"):
        assert callable(getattr(ucv, name))
    for name in ("split_to_batches", "positional_encoding_for_views", "positional_encoding_for_xyz", "ray_marching",
                 "get_psnr", "get_psnr_for_image", "prepare_ds", "c2w_to_rays_prepare_ds", "render_rays",
                 "model_predict", "get_num_of_batches"):
        assert callable(getattr(unrf, name))
    for name in ("render", "call", "train_step", "render_rays", "render_image", "init_network", "get_nerf_model_path"):
        assert hasattr(pkg.NeRFModel, name)
    assert pkg.DietNeRFModel.COARSE_LOSS_WEIGHT == 2.0
    assert str(pkg.NeRFModel.get_nerf_model_path("x", 7)).endswith("saved_weights/NeRF_model_epoch_007.h5")
    assert unrf.get_num_of_batches(4096, 70, 50, 50) == 42
    assert unrf.get_size_of_splits(4096, 10000) == [4096, 4096, 1808]
    assert unrf.get_size_of_splits(4096, 2500) == [2500]
    assert unrf.get_size_of_splits(100, 200) == [100, 100]


def test_missing_library_fails_loudly(pkg, monkeypatch):
    monkeypatch.setattr(pkg._lib, "_lib", None)
    monkeypatch.setattr(pkg._lib, "LIB_PATH", "/nonexistent/libnerf_b200.so")
    with pytest.raises(pkg.NerfLibraryError):
        pkg._lib.load()


def test_hierarchical_sample_reports_argument_errors(pkg):
    """nerf_hierarchical_sample (the step between the two networks of NeRF.render, src/NeRF.py:129-132, as one launch):
    its argument checks need no GPU."""
    lib = pkg.load()
    p256 = ctypes.c_void_p(256)
    args = lambda raw, z, n, s, nf, out: (raw, z, n, s, nf, 1, 0, 0, out, None)
    assert lib.nerf_hierarchical_sample(*args(None, p256, 4, 64, 128, p256)) == -1 and b"null pointer" in lib.nerf_last_error()
    assert lib.nerf_hierarchical_sample(*args(p256, p256, 4, 64, 300, p256)) == -1 and b"n_new <= 256" in lib.nerf_last_error()
    assert lib.nerf_hierarchical_sample(*args(p256, p256, 4, 1, 128, p256)) == -1
    assert lib.nerf_hierarchical_sample(*args(ctypes.c_void_p(260), p256, 4, 64, 128, p256)) == -1
    assert b"16-byte aligned" in lib.nerf_last_error()
    assert lib.nerf_hierarchical_sample(*args(p256, p256, 0, 64, 128, p256)) == 0       # an empty ray range is not an error


def test_fused_entry_points_report_argument_errors(pkg):
    """nerf_render_fused_fwd / nerf_train_step_fused: workspace queries and argument checks need no GPU."""
    lib = pkg.load()
    L = pkg._lib
    cfg = pkg.NetCfg(5, 4, 2, 256, 128, 0.05)
    rc = L.RenderCfg(0.5576, 2.5635, 64, 128, L.MODE_BF16)
    n_bytes = lib.nerf_render_workspace_bytes(ctypes.byref(cfg), ctypes.byref(rc), 1000)
    t_bytes = lib.nerf_train_workspace_bytes(ctypes.byref(cfg), ctypes.byref(rc), 1000)
    assert n_bytes >= 1000 * (64 + 192 * 4 + 64 + 128 + 192) * 4 and t_bytes > n_bytes
    assert lib.nerf_render_workspace_bytes(ctypes.byref(cfg), ctypes.byref(L.RenderCfg(0.5576, 2.5635, 0, 128, 1)), 10) == -1
    # the fp16-operand mode trains too (round 2): same workspace as the bf16 mode, whose backward it shares
    assert lib.nerf_train_workspace_bytes(ctypes.byref(cfg), ctypes.byref(L.RenderCfg(0.5576, 2.5635, 64, 128, L.MODE_FP16)), 1000) == t_bytes
    assert lib.nerf_train_workspace_bytes(ctypes.byref(cfg), ctypes.byref(L.RenderCfg(0.5576, 2.5635, 64, 128, 7)), 10) == -1
    rng, outs = L.RngState(1, 0, 0, 0), L.RenderOuts()
    st = lib.nerf_render_fused_fwd(ctypes.byref(cfg), ctypes.byref(rc), None, None, None, None, ctypes.c_void_p(256),
                                   ctypes.c_void_p(256), 4, ctypes.byref(rng), ctypes.byref(outs), ctypes.c_void_p(256), None)
    assert st == -1 and b"weights missing" in lib.nerf_last_error()
    st = lib.nerf_render_fused_fwd(ctypes.byref(cfg), ctypes.byref(rc), None, ctypes.c_void_p(256), None, ctypes.c_void_p(256),
                                   ctypes.c_void_p(256), ctypes.c_void_p(256), 4, ctypes.byref(rng), ctypes.byref(outs),
                                   ctypes.c_void_p(264), None)
    assert st == -1 and b"256-byte aligned" in lib.nerf_last_error()
    # the xyz-only network (n_angles_for_model: 0, src/NeRF.py:248-288) has a tensor-core plan as well (round 2); a network
    # outside the plans (hidden width != 256) still says so and sizes only the fp32 mode
    cfg0 = pkg.NetCfg(5, 4, 0, 256, 128, 0.05)
    assert lib.nerf_packed_bytes(ctypes.byref(cfg0)) > 0
    assert lib.nerf_render_workspace_bytes(ctypes.byref(cfg0), ctypes.byref(rc), 10) > 0
    assert lib.nerf_train_workspace_bytes(ctypes.byref(cfg0), ctypes.byref(rc), 10) > 0
    cfg1 = pkg.NetCfg(5, 4, 2, 128, 128, 0.05)
    assert lib.nerf_render_workspace_bytes(ctypes.byref(cfg1), ctypes.byref(rc), 10) == -1
    assert lib.nerf_train_workspace_bytes(ctypes.byref(cfg1), ctypes.byref(rc), 10) == -1
    assert lib.nerf_render_workspace_bytes(ctypes.byref(cfg1), ctypes.byref(L.RenderCfg(0.5576, 2.5635, 64, 128, L.MODE_FP32)), 10) > 0
    st = lib.nerf_render_fused_fwd(ctypes.byref(cfg1), ctypes.byref(rc), None, ctypes.c_void_p(256), None, ctypes.c_void_p(256),
                                   ctypes.c_void_p(256), ctypes.c_void_p(256), 4, ctypes.byref(rng), ctypes.byref(outs),
                                   ctypes.c_void_p(256), None)
    assert st == -3 and b"no tensor-core path" in lib.nerf_last_error()
    # the sharded one-call step: a peer exchange is mandatory, and it needs the Adam state and a rank inside the world
    tcfg = L.TrainCfg(1.0, 0, 0, 5e-4, 0.9, 0.999, 1e-7)
    p256 = ctypes.c_void_p(256)
    common = (ctypes.byref(cfg), ctypes.byref(rc), ctypes.byref(tcfg), p256, p256, p256, p256, p256, p256, p256, 4, 8,
              ctypes.byref(rng), p256)
    st = lib.nerf_train_step_fused_sharded(*common, p256, p256, 1, None, p256, None, None, None)
    assert st == -1 and b"null peer exchange" in lib.nerf_last_error()
    peer = L.PeerExchangeCfg(256, 256, 256, 2, 2, 1, 0)                 # rank 2 of a world of 2
    st = lib.nerf_train_step_fused_sharded(*common, p256, p256, 1, None, p256, None, ctypes.byref(peer), None)
    assert st == -1 and b"valid rank" in lib.nerf_last_error()
    peer = L.PeerExchangeCfg(256, 256, 256, 0, 2, 1, 0)
    st = lib.nerf_train_step_fused_sharded(*common, None, None, 1, None, p256, None, ctypes.byref(peer), None)
    assert st == -1 and b"Adam state" in lib.nerf_last_error()


def test_c_caller_links_and_fails_loudly_without_a_gpu(pkg):
    """tools/c_caller_demo.c (plain C, no Python, no torch) builds against include/nerf_b200.h + libnerf_b200.so; without
    a CUDA device it stops at the first CUDA call with a message and a non-zero exit code, never with made-up numbers."""
    import subprocess
    import torch
    import __graft_entry__ as entry
    if not os.path.exists(entry.C_DEMO):
        entry.build()
    assert os.path.exists(entry.C_DEMO)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is test_c_caller_trains_and_renders")
    r = subprocess.run([entry.C_DEMO, "3"], capture_output=True, text=True, timeout=60)
    assert r.returncode in (2, 3) and r.stdout.strip() == "" and ("CUDA" in r.stderr or "cuda" in r.stderr)


def test_ctypes_structs_match_the_header(pkg, tmp_path):
    """sizeof / offsetof of every struct of include/nerf_b200.h as gcc lays them out == the ctypes mirrors in _lib.py."""
    import subprocess
    L = pkg._lib
    structs = {"nerf_net_cfg": L.NetCfg, "nerf_render_cfg": L.RenderCfg, "nerf_rng_state": L.RngState,
               "nerf_render_outs": L.RenderOuts, "nerf_train_cfg": L.TrainCfg, "nerf_peer_exchange": L.PeerExchangeCfg}
    lines = []
    for cname, mirror in structs.items():
        lines.append(f'printf("{cname} %zu", sizeof({cname}));')
        for field, _ in mirror._fields_:
            lines.append(f'printf(" %zu", offsetof({cname}, {field}));')
        lines.append('printf("\\n");')
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "nerf_b200.h"\nint main(void) {\n' + "\n".join(lines) +
                   "\nreturn 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(structs)
    for line in out:
        name, size, *offsets = line.split()
        mirror = structs[name]
        assert ctypes.sizeof(mirror) == int(size), name
        assert [getattr(mirror, f).offset for f, _ in mirror._fields_] == [int(o) for o in offsets], name


def test_mlp_kernels_are_tcgen05_kernels(pkg):
    """Static check of the built library (cuobjdump, no GPU): the three MLP kernels issue tcgen05.mma (SASS UTC*MMA), read
    TMEM (LDTM) and move operands with the TMA engine's bulk copies (UBLKCP); no kernel falls back to legacy mma.sync
    (HMMA) -- the guide's 'what proves a Blackwell-native kernel'."""
    import re
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = []
        elif cur and "/*" in line:
            per_kernel[cur].append(line)
    assert len(per_kernel) >= 40
    assert not any(re.search(r"\bHMMA", l) for body in per_kernel.values() for l in body)
    for stem in ("mlp_tc_fwd_kernel", "mlp_tc_bwd_chain_kernel", "mlp_tc_bwd_dw_kernel"):
        bodies = [b for k, b in per_kernel.items() if stem in k]
        assert bodies, stem
        for body in bodies:
            text = "\n".join(body)
            assert re.search(r"\bUTC[A-Z]*MMA", text) and "LDTM" in text and "UBLKCP" in text, stem
