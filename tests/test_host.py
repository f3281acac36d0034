"""CPU: host-side logic of the product package -- the C-ABI library loads and exports every symbol the
header declares, the reference-compatible modules accept the reference's state dict, the folded
weights reproduce the oracle (through a test-only emulation of the kernel algebra), CPU tensors are
rejected loudly, and the N>1 sharding/all-gather logic works over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden, relerr, state_dict_for


def test_library_exports_every_declared_symbol():
    import ctypes

    from livingscenes_b200 import _lib

    header = open(os.path.join(ROOT, "include", "livingscenes_b200.h")).read()
    declared = set(re.findall(r"LS_API\s+(?:const\s+char\*|int64_t|int)\s+(ls_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name)
    assert handle.ls_version() == 3
    # argument validation happens before any CUDA call
    n = ctypes.c_size_t(0)
    assert handle.ls_encoder_workspace_bytes(None, 1, 1024, ctypes.byref(n)) == -1
    assert b"null encoder descriptor" in handle.ls_last_error()


def test_struct_layout_matches_header():
    import ctypes

    from livingscenes_b200 import _lib

    assert ctypes.sizeof(_lib.EncLayerDesc) == 88
    assert ctypes.sizeof(_lib.EncoderDesc) == 24 + 8 * 88 + 5 * 8 + 8 + 8
    assert ctypes.sizeof(_lib.EncoderIO) == 8 + 16 + 5 * 8 + 3 * 64 + 16 + 2 * 64
    assert ctypes.sizeof(_lib.DecoderDesc) == 16 + 2 * 96 + 16 + 2 * 48 + 96 + 2 * 96 + 4 * 8


def test_state_dict_keys_match_reference_checkpoint():
    import livingscenes_b200 as ls
    from oracle import restatement as R

    sd = R.random_state_dict(0)
    sp = ls.Shape_Prior.from_state_dict(sd)
    enc_keys = {k[len("encoder."):] for k in sd if k.startswith("encoder.")}
    dec_keys = {k[len("decoder."):] for k in sd if k.startswith("decoder.")}
    assert set(sp.encoder.state_dict().keys()) == enc_keys
    assert set(sp.decoder.F.state_dict().keys()) == dec_keys
    assert sum(p.numel() for p in sp.parameters()) == 7395808


def test_unsupported_configurations_fail_loudly():
    import livingscenes_b200 as ls

    with pytest.raises(NotImplementedError):
        ls.VecDGCNN_att(num_knn=8)
    with pytest.raises(NotImplementedError):
        ls.VecDGCNN_att(use_dg=False)
    with pytest.raises(NotImplementedError):
        ls.VecDGCNN_att(atten_multi_head_c=8, feat_dim=[32] * 8)


def test_cpu_tensors_are_rejected_no_fallback():
    import livingscenes_b200 as ls
    from oracle import restatement as R

    sp = ls.Shape_Prior.from_state_dict(R.random_state_dict(0))
    with pytest.raises(RuntimeError, match="no CPU path"):
        sp.encode(torch.zeros(2, 3, 1024))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ls.sequential_matcher(torch.zeros(3, 256), torch.zeros(3, 256))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ls.kabsch_transformation_estimation(torch.zeros(1, 8, 3), torch.zeros(1, 8, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "livingscenes_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


@pytest.mark.parametrize("tag", ["random", "shipped"])
def test_folded_weights_reproduce_golden(tag):
    import livingscenes_b200 as ls
    from emulate import encoder_emulated

    sd = state_dict_for(tag)
    g = golden(f"encoder_{tag}")
    sp = ls.Shape_Prior.from_state_dict(sd)
    xn = torch.from_numpy(g["x_norm"])[:1]
    knn = [torch.from_numpy(g[f"knn_idx_{i}"][:1].astype(np.int64)) for i in range(7)]
    fps = [torch.from_numpy(g[f"fps_idx_{i}"][:1].astype(np.int64)) for i in range(3)]
    with torch.no_grad():
        c, s, zs, zi, feats = encoder_emulated(sp.encoder, xn, knn, fps)
    assert relerr(c, g["center"][:1]) < 1e-4
    assert relerr(s, g["scale"][:1]) < 1e-4
    assert relerr(zs, g["z_so3"][:1]) < 1e-4
    assert relerr(zi, g["z_inv"][:1]) < 1e-4
    for i, f in enumerate(feats):
        assert relerr(f[..., ::16], g[f"feat_{i}"][:1]) < 1e-4, i


def test_decoder_packing_is_consistent():
    """The collapsed layer-0 / layer-4 weights reproduce the 513-wide formulation."""
    import livingscenes_b200 as ls
    from oracle import restatement as R

    sd = R.random_state_dict(0)
    dec = ls.Shape_Prior.from_state_dict(sd).decoder.F
    Ws = R.decoder_weights(sd)
    L = 256
    eff = [getattr(dec, f"lin{l}").effective() for l in range(9)]
    for l in range(9):
        assert torch.allclose(eff[l][0], Ws[l][0], atol=0, rtol=1e-6)
    u = torch.randn(513)
    h3 = torch.randn(255)
    W0, W4 = eff[0][0], eff[4][0]
    full0 = W0 @ u
    split0 = W0[:, :L] @ u[:L] + W0[:, L:] @ u[L:]
    assert torch.allclose(full0, split0, atol=1e-5)
    full4 = W4 @ torch.cat([h3, u])
    split4 = torch.cat([W4[:, :255], W4[:, 255 + L:]], 1) @ torch.cat([h3, u[L:]]) + W4[:, 255:255 + L] @ u[:L]
    assert torch.allclose(full4, split4, atol=1e-5)


def test_tc_packed_floats_accounts_for_both_tile_formats():
    """ls_tc_packed_floats (host-only entry point): the packed buffer holds the 128-row-tile images of k_gemm_tc /
    k_gemm_tc2 and, for R > 128, the 256-row-tile images of k_gemm_tc3 behind them (hi + lo, k-blocks of 16)."""
    import ctypes as C

    from livingscenes_b200 import _lib

    def packed(R, K):
        n = C.c_size_t(0)
        assert _lib.lib().ls_tc_packed_floats(R, K, C.byref(n)) == 0
        return n.value

    ceil = lambda a, b: -(-a // b)
    for R, K in ((64, 32), (128, 64), (129, 64), (256, 32), (257, 512), (768, 768), (2048, 256), (1, 1)):
        kb = ceil(K, 16)
        want = ceil(R, 128) * kb * 2 * 128 * 16 + (ceil(R, 256) * kb * 2 * 256 * 16 if R > 128 else 0)
        assert packed(R, K) == want, (R, K)
    n = C.c_size_t(0)
    assert _lib.lib().ls_tc_packed_floats(0, 8, C.byref(n)) != 0  # invalid sizes are rejected, not clamped


def test_roofline_traffic_file_is_reproducible_from_the_committed_launch_list(tmp_path):
    """profiles/knn_edgeconv_traffic.json (bench.py's roofline.traffic) is exactly what scripts/summarize_launches.py
    derives from the committed ncu launch list of the final commit."""
    import json

    out = tmp_path / "traffic.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"),
                        os.path.join(ROOT, "profiles", "r02", "launches_fwd.csv"), "--traffic", str(out)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got, want = json.load(open(out)), json.load(open(os.path.join(ROOT, "profiles", "knn_edgeconv_traffic.json")))
    for layer in range(7):
        k = f"layer{layer}_dram_bytes_per_launch"
        assert got[k] == want[k] and got[k] > 0, k
    assert "k_knn_edge" in r.stdout and "k_gemm_tc3" in r.stdout


def test_shard_ranges_cover_and_are_contiguous():
    from livingscenes_b200.dist import shard_range

    for n in (0, 1, 7, 128, 129):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    from livingscenes_b200.dist import pack_codes, unpack_codes

    code = {"z_so3": torch.randn(5, 256, 3), "z_inv": torch.randn(5, 256), "s": torch.rand(5), "t": torch.randn(5, 1, 3)}
    rec = pack_codes(code)
    assert rec.shape == (5, 1028)
    back = unpack_codes(rec)
    for k in code:
        assert torch.equal(back[k], code[k])


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from livingscenes_b200.dist import all_gather_codes, shard_range, CODE_FLOATS
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
n_total = int(sys.argv[4])
full = torch.arange(n_total * CODE_FLOATS, dtype=torch.float32).reshape(n_total, CODE_FLOATS)
lo, hi = shard_range(n_total, dist.get_rank(), 2)
out = all_gather_codes(full[lo:hi].clone(), n_total)
assert out.shape == full.shape and torch.equal(out, full), "all-gather mismatch"
dist.barrier()
dist.destroy_process_group()
print("ok")
"""


@pytest.mark.parametrize("n_total", [8, 7])
def test_all_gather_codes_gloo_world2(tmp_path, n_total):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29600 + os.getpid() % 300 + n_total)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), str(n_total)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o
