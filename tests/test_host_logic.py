"""CPU: host-side logic — shape algebra, parameter naming contract, the C-ABI library's exports,
the product/oracle separation, and the multi-rank batch sharding (gloo, world_size 2)."""
import os
import re
import socket
import subprocess
import sys

import pytest
import torch

import cse_b200  # noqa: F401
from cse_b200 import _lib, shapes, sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shape_algebra_matches_survey_table():
    for T, L, S in [(16000, 1999, 18), (32000, 3999, 34), (64000, 7999, 66), (128000, 15999, 130),
                    (240000, 29999, 242), (256000, 31999, 258)]:
        ps = shapes.path_shape(1, T, 1, 2)
        assert (ps.L, ps.S, ps.gap, ps.T_est) == (L, S, 126, T)
    assert abs(shapes.algorithmic_flops(shapes.path_shape(1, 32000, 0, 2)) / 1e9 - 473.4) < 0.1
    assert abs(shapes.algorithmic_flops(shapes.path_shape(16, 32000, 1, 2)) / 1e9 - 7697.3) < 0.1
    with pytest.raises(ValueError):
        shapes.path_shape(1, 15)


def test_c_abi_shape_function_agrees_with_python():
    for B, T, c, spk in [(1, 16, 0, 2), (2, 4003, 1, 2), (16, 32000, 1, 2), (3, 128000, 2, 3)]:
        s = _lib.path_shape(B, T, c, spk)
        ps = shapes.path_shape(B, T, c, spk)
        assert (s.L, s.gap, s.S, s.T_est) == (ps.L, ps.gap, ps.S, ps.T_est)
    with pytest.raises(_lib.CseError):
        _lib.path_shape(1, 8, 0, 2)
    assert _lib.load().cse_workspace_bytes(16, 32000, 1, 2, _lib.BF16) > 0


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cse_b200.h")).read()
    declared = set(re.findall(r"CSE_API\s+[\w\s\*]+?\b(cse_\w+)\s*\(", header))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cse_b200.h but not exported"
    assert declared == set(_lib.exported_names()), declared ^ set(_lib.exported_names())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "contextual-speech-extraction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "sb_shim" not in text, f


def test_cuda_path_fails_loudly_without_gpu_tensors():
    from cse_b200.models.sepformer import Sepformer
    m = Sepformer(2)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4000))
    from cse_b200 import losses
    with pytest.raises(RuntimeError):
        losses.get_si_snr_with_pitwrapper(torch.zeros(1, 100, 2), torch.zeros(1, 100, 2))


def test_state_dict_contract_all_variants():
    from cse_b200.models.ContExt import Sepformer as ContExt
    from cse_b200.models.ContSep import Sepformer as ContSep
    from cse_b200.models.sepformer import Sepformer as Plain
    m = ContSep(3, add_mt=True, ce=True)
    m.add_mt_pipeline()
    assert set(m.state_dict()) == set(synth.param_spec("contsep", 3))
    m = ContSep(2, add_mt=True, ce=False)
    m.add_mt_pipeline()
    assert m.context_selector.weight.shape == (1, 256)
    m = Plain(2)
    assert set(m.state_dict()) == set(synth.param_spec("sepformer", 2))
    m = ContExt(2, add_ctx=True, add_se=True)
    m.add_ctx_pipeline()
    m.add_se_pipeline()
    spec = synth.param_spec("hcontext", 2)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: s for k, (s, _) in spec.items()}
    assert sum(p.numel() for p in m.parameters()) == 30665217          # SURVEY.md §8 (HContExt 2-spk)


def test_shard_planners():
    assert [list(sharding.contiguous_shard(16, r, 8)) for r in range(8)] == [[2 * r, 2 * r + 1] for r in range(8)]
    cover = sum((list(sharding.contiguous_shard(5, r, 4)) for r in range(4)), [])
    assert cover == list(range(5))
    assert len(sharding.contiguous_shard(1, 3, 4)) == 0                  # idle rank when B < world
    lens = [8, 1, 7, 2, 6, 3, 5, 4]
    sh = sharding.balanced_shards(lens, 2)
    assert sorted(sum(sh, [])) == list(range(8))
    loads = [sum(lens[i] for i in s) for s in sh]
    assert abs(loads[0] - loads[1]) <= 1


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import cse_b200
from cse_b200 import sharding, synth
from oracle import sepformer_oracle as O     # test-only stand-in for the CUDA model on CPU
rank, world, port = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
sd = {"encoder.conv1d.weight": synth.make_state_dict("sepformer", 2, seed=3)["encoder.conv1d.weight"],
      "decoder.weight": synth.make_state_dict("sepformer", 2, seed=3)["decoder.weight"]}
def model_fn(mix, ctx):
    w = O.encoder(sd, mix)
    return torch.stack([O.decoder(sd, w), O.decoder(sd, 0.5 * w)], -1)[:, : mix.shape[1]]
mix, _ = synth.make_mixture(5, 816, 2, seed=9)
full = model_fn(mix, None)
got = sharding.separate_sharded(model_fn, mix, None)
assert got.shape == full.shape, (got.shape, full.shape)
assert torch.allclose(got, full, atol=1e-6)
local = sharding.separate_sharded(model_fn, mix, None, gather=False)
assert local.shape[0] == len(sharding.contiguous_shard(5, rank, world))
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_sharded_separation_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_training_entry_points_validate_arguments_before_any_launch():
    """The backward / training C-ABI entry points reject bad shapes and NULL pointers with a
    message (no CUDA call is reached, so this runs without a GPU)."""
    import ctypes as C
    lib = _lib.load()
    p = C.c_void_p(0x1000)                       # never dereferenced: every call below fails validation first
    n = 4 * 36
    per_row = 256 + 768 + 256 + 256 + 1024 + 1024 + 256
    assert lib.cse_layer_workspace_bytes(4, 36) == n * per_row * 4 + 1024 * 256 * 4
    assert lib.cse_layer_workspace_bytes(0, 36) == 0
    cases = [
        ("cse_linear_bwd", (p, 256, p, p, 300, 10, 300, 256, None, 256, p, None, None, None), "multiples of 128"),
        ("cse_linear_bwd", (None, 256, p, p, 256, 10, 256, 256, None, 256, p, None, None, None), "NULL"),
        ("cse_attention_bwd", (p, p, p, 1, 1000, p, None), "does not fit shared memory"),
        ("cse_gate_bwd", (p, p, p, 12, p, p, None), "multiple of 8"),
        ("cse_mask_decode_bwd", (p, p, p, p, 1, 10, 100, 5, p, p, None, None), "n_masks=5"),
        ("cse_encoder_bwd", (p, p, p, 1, 8, p, None), "shorter than the encoder kernel"),
        ("cse_groupnorm_bwd", (p, p, p, p, 0, 10, p, None, None, p, None), "bad argument"),
        ("cse_sequences_to_chunks", (p, 1, 4, 1, 0, None, None, None), "bad argument"),
        ("cse_si_snr_bwd", (p, p, None, 1, 100, 2, p, p, None), "bad argument"),
        ("cse_layer_bwd", (None, None, p, p, 1, 10, p, 0, None), "bad argument"),
    ]
    for name, args, msg in cases:
        with pytest.raises(_lib.CseError, match=msg):
            _lib.call(name, *args)
    lp, lg = _lib.LayerParams(), _lib.LayerGrads()
    with pytest.raises(_lib.CseError, match="workspace too small"):
        _lib.call("cse_layer_bwd", C.byref(lp), C.byref(lg), p, p, 4, 36, C.c_void_p(0x10000), 16, None)
    with pytest.raises(_lib.CseError, match="256-byte aligned"):
        _lib.call("cse_layer_fwd", C.byref(lp), p, 4, 36, 0, C.c_void_p(0x10010), 1 << 30, None)


def test_losses_and_training_refuse_cpu_tensors():
    from cse_b200 import losses, training
    a = torch.zeros(2, 100, 2)
    with pytest.raises(_lib.CseError):
        losses.get_si_snr_with_pitwrapper(a, a)
    with pytest.raises(_lib.CseError):
        training.forward_train({}, torch.zeros(1, 4000), None, 2)


def test_experimental_tensor_core_backward_entry_validates_arguments():
    import ctypes as C
    lib = _lib.load()
    assert lib.cse_linear_bwd_tc_scratch_bytes(100, 256, 256) >= 100 * 256 * 2 + 256 * 256 * 6 + 2 * 256 * 128 * 2
    assert lib.cse_linear_bwd_tc_scratch_bytes(0, 256, 256) == 0
    p = C.c_void_p(0x1000)
    with pytest.raises(_lib.CseError, match="multiples of 128"):
        _lib.call("cse_linear_bwd_tc", p, 0, 256, p, p, 10, 200, 256, p, 1, 256, p, None, p, 1 << 30, None)
    with pytest.raises(_lib.CseError, match="scratch too small"):
        _lib.call("cse_linear_bwd_tc", p, 0, 256, p, p, 10, 256, 256, p, 1, 256, p, None, p, 16, None)
