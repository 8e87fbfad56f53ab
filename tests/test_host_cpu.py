"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
the host modules mirror the reference's constructors / attributes / state_dict layout, and the
product path refuses to run without CUDA instead of falling back."""
import re
from pathlib import Path

import pytest
import torch

import quantizedsae_b200 as Q
from quantizedsae_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def _header_functions():
    text = (ROOT / "include" / "qsae_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsae_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()                      # builds with nvcc if missing; raises if it cannot
    declared = _header_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/qsae_b200.h but not exported"
    assert set(declared) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.qsae_abi_version() == 1


def test_default_sample_rows_and_its_tuning_switch(monkeypatch):
    """Rows of the stratified encoder sample (host logic of the sampled-prior path, no device call): H / 16 in whole
    256-row tiles for dictionaries of at least 8192 latents, none below; QSAE_SAMPLE_DIV is read once and re-read by
    qsae_reload_tuning; the workspace query accepts the recommended size."""
    import ctypes as C

    lib = _lib.load()
    monkeypatch.delenv("QSAE_SAMPLE_DIV", raising=False)
    assert lib.qsae_reload_tuning() == 0
    assert [_lib.default_sample_rows(h) for h in (1024, 8191, 8192, 32768, 2 ** 17, 2 ** 20)] == [0, 0, 512, 2048, 8192, 65536]
    assert _lib.default_sample_rows(40000) == 2560        # 2500 rounded up to whole tiles
    monkeypatch.setenv("QSAE_SAMPLE_DIV", "32")
    assert _lib.default_sample_rows(32768) == 2048        # cached until reloaded
    assert lib.qsae_reload_tuning() == 0
    assert _lib.default_sample_rows(32768) == 1024
    monkeypatch.setenv("QSAE_SAMPLE_DIV", "2")             # clamped: the plan samples only when H >= 8 n_sample
    assert lib.qsae_reload_tuning() == 0
    assert _lib.default_sample_rows(32768) == 4096
    monkeypatch.delenv("QSAE_SAMPLE_DIV")
    assert lib.qsae_reload_tuning() == 0
    n = C.c_size_t(0)
    assert lib.qsae_encode_topk_workspace_bytes(4096, 32768, 512, 32, _lib.default_sample_rows(32768), C.byref(n)) == 0
    assert n.value > 0


def test_bench_reference_arm_line_on_cpu():
    """`bench.py --impl reference` needs no GPU: it times the op-for-op CPU port of the reference forward and prints the
    contract's JSON line (same metric / unit as the B200 arm, impl = reference, cpu_baseline and a zero-copy e2e block)."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-batch", "32"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["higher_is_better"] is True
    assert line["metric"] == "b_sae 512->32768 4-bit fwd tokens/s" and line["value"] > 0 and line["steps"] == 1
    assert line["config"]["cpu_rows_per_step"] == 32 and "32 rows" in line["config"]["note"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "32 rows" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_sass_is_blackwell_native():
    """tcgen05 / TMA / TMEM loads must be present in the built library (no GPU needed)."""
    import shutil
    import subprocess

    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists():
        pytest.skip("cuobjdump not available")
    _lib.load()
    sass = subprocess.run([exe, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from libqsae_b200.so SASS"
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync path found"


def test_argument_validation_without_gpu():
    """Pure host-side validation paths of the C ABI (they return before touching a device)."""
    import ctypes as C

    lib = _lib.load()
    n = C.c_size_t(0)
    assert lib.qsae_encode_topk_workspace_bytes(128, 1024, 500, 8, 0, C.byref(n)) == -1   # D % 8
    assert b"multiple of 8" in lib.qsae_last_error()
    assert lib.qsae_encode_topk_workspace_bytes(128, 1024, 1024, 8, 0, C.byref(n)) == -1  # D > 512
    assert lib.qsae_encode_topk_workspace_bytes(128, 16, 64, 32, 0, C.byref(n)) == -5     # k > H
    assert b"out of range" in lib.qsae_last_error()
    assert lib.qsae_encode_topk_workspace_bytes(128, 8192, 64, 5000, 0, C.byref(n)) == -1  # k > MAX_K_LARGE
    assert b"QSAE_MAX_K_LARGE" in lib.qsae_last_error()
    assert lib.qsae_encode_topk_workspace_bytes(128, 4096, 64, 500, 0, C.byref(n)) == 0 and n.value > 128 * 4096 * 4  # dense path
    assert lib.qsae_encode_topk_workspace_bytes(4096, 131072, 512, 262, 4096, C.byref(n)) == 0 and n.value > 0       # prior path
    assert lib.qsae_encode_topk_workspace_bytes(4096, 32768, 512, 32, 1024, C.byref(n)) == 0 and n.value > 0
    assert lib.qsae_tsae_workspace_bytes(128, 1004, 64, 0, C.byref(n)) == -1             # H % 8
    assert lib.qsae_tsae_workspace_bytes(4096, 32768, 512, 1, C.byref(n)) == 0 and n.value > 2 * 4096 * 32768 * 2
    assert lib.qsae_decode_dense_workspace_bytes(128, 1024, 514, C.byref(n)) == -1       # N % 4
    assert lib.qsae_pack_bitplanes(None, 8, 8, 4, None, None, None) == -1
    assert lib.qsae_decode_int4(None, None, 1, 1, None, 8, 8, 1.0, None, None, None) == -1
    with pytest.raises(RuntimeError):
        _lib.check(-5)
    with pytest.raises(_lib.QsaeError):
        _lib.check(-1)


def test_bsae_constructor_attributes_and_state_dict():
    m = Q.BinarySAE(64, 2048, 4.0, 4)                     # positional, like scripts/training/train.py:60-91
    assert (m.input_dim, m.hidden_dim, m.n_bits, m.k) == (64, 2048, 4, 0.002)
    assert isinstance(m.encoder, torch.nn.Sequential) and isinstance(m.encoder[0], torch.nn.Linear)
    d = m.decoder
    assert (d.in_features, d.out_features, d.n_bits, d.gamma, d.quantization_step) == (2048, 64, 4, 4.0, 0.5)
    assert d.scale_factor == 16
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "encoder.0.weight": (2048, 64), "encoder.0.bias": (2048,),
        "decoder.weight": (2048, 256), "decoder.bias": (64,)}
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert float(m.encoder[0].bias.abs().max()) == 0.0     # zeros_ init (sae/binary.py:87)
    bound = (6.0 / (64 + 2048)) ** 0.5                      # xavier_uniform gain 1 (sae/binary.py:86)
    assert float(m.encoder[0].weight.abs().max()) <= bound + 1e-6
    m8 = Q.BinarySAE(32, 1024)                              # defaults gamma=4.0, n_bits=8
    assert m8.n_bits == 8 and m8.decoder.quantization_step == 4.0 / 128


def test_bsae_state_dict_matches_reference_fixture(golden_dir):
    import numpy as np

    g = np.load(golden_dir / "bsae_polar_d64_h2048.npz")
    m = Q.BinarySAE(64, 2048, 4.0, 4)
    assert sorted(m.state_dict().keys()) == g["state_keys"].tolist()
    assert [str(tuple(v.shape)) for _, v in sorted(m.state_dict().items())] == g["state_shapes"].tolist()
    g = np.load(golden_dir / "baseline_d64_h2048.npz")
    b = Q.BaselineSparseAutoencoder(64, 2048)
    assert sorted(b.state_dict().keys()) == g["state_keys"].tolist()
    assert [str(tuple(v.shape)) for _, v in sorted(b.state_dict().items())] == g["state_shapes"].tolist()
    assert b.topk == 32


def test_qsae_constructor_and_state_dict(golden_dir):
    import numpy as np

    g = np.load(golden_dir / "qsae_d64_h2048.npz")
    m = Q.QuantizedMatryoshkaSAE(64, 2048, 32, 4.0, 4)      # (input_dim, hidden_dim, top_k, abs_range, n_bits)
    assert sorted(m.state_dict().keys()) == g["state_keys"].tolist()
    assert [str(tuple(v.shape)) for _, v in sorted(m.state_dict().items())] == g["state_shapes"].tolist()
    assert m.decoder.nested_dictionary_size == g["level_sizes"].tolist() == [256, 256, 512, 1024]
    assert (m.top_k, m.abs_range, m.n_bits, m.allow_bias, m.decoder.quant_step) == (32, 4.0, 4, True, 0.5)
    assert isinstance(m.encoder[1], torch.nn.Sigmoid)
    assert Q.QuantizedMatryoshkaSAE(512, 32768, 32, 4, 4).decoder.nested_dictionary_size == [4096, 4096, 8192, 16384]
    d = Q.QuantizedMatryoshkaDecoder(1024, 64)               # defaults abs_range=4, n_bits=8
    assert d.n_bits == 8 and d.quant_step == 4 / 128 and sum(d.nested_dictionary_size) == 1024


def test_tsae_constructor_and_state_dict(golden_dir):
    import numpy as np

    g = np.load(golden_dir / "tsae_d64_h2048.npz")
    m = Q.TernarySparseAutoencoder(64, 2048)
    assert sorted(m.state_dict().keys()) == g["state_keys"].tolist()
    assert [str(tuple(m.state_dict()[k].shape)) for k in sorted(m.state_dict())] == g["state_shapes"].tolist()
    assert m.topk == 4 and m.decoder.threshold == 0.5
    assert isinstance(m.encoder[1], torch.nn.ReLU) and tuple(m.decoder.weight.shape) == (64, 2048)
    assert float(m.decoder.mask.min()) == 1.0 and m.decoder.input_activations is None
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(4, 64))


def test_state_dict_round_trip_strict():
    a, b = Q.BinarySAE(64, 1024, 1.5, 4), Q.BinarySAE(64, 1024, 1.5, 4)
    b.load_state_dict(a.state_dict(), strict=True)
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k])


def test_base_class_contract():
    s = Q.SparseAutoencoder(8, 16)
    with pytest.raises(NotImplementedError):
        s.encode(torch.zeros(1, 8))
    with pytest.raises(NotImplementedError):
        s.decode(torch.zeros(1, 16))


def test_no_cpu_fallback():
    m = Q.BinarySAE(64, 1024, 4.0, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(4, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Q.BaselineSparseAutoencoder(64, 1024)(torch.zeros(4, 64))
    with pytest.raises(_lib.QsaeError):
        _lib.cast_bf16(torch.zeros(8))


def test_prepared_cache_invalidation():
    from quantizedsae_b200.sae.base import PreparedCache, param_key

    p = torch.nn.Parameter(torch.zeros(4))
    cache, calls = PreparedCache(), []
    make = lambda: calls.append(1) or len(calls)
    assert cache.get("x", param_key(p), make) == 1
    assert cache.get("x", param_key(p), make) == 1          # cached
    with torch.no_grad():
        p.add_(1.0)                                         # in-place update bumps _version
    assert cache.get("x", param_key(p), make) == 2
    p.data = torch.ones(4)                                  # load_state_dict / .to() style swap
    assert cache.get("x", param_key(p), make) == 3


def test_oracle_is_not_imported_by_the_product():
    """The package must never reach into oracle/ (test infrastructure only)."""
    for path in (ROOT / "quantizedsae_b200").rglob("*.py"):
        text = path.read_text()
        assert "oracle" not in re.sub(r"#.*", "", text).replace("no oracle", ""), path


def test_inference_registry_mirrors_the_reference(tmp_path):
    """Registry keys / kwargs of inference/framework.py:165-220, checkpoint restore, key remap."""
    from quantizedsae_b200 import inference as I
    from quantizedsae_b200.inference.framework import remap_eleuther_keys

    assert {"b_sae", "q_sae", "rq_sae", "baseline_sae"} <= set(I.SAE_REGISTRY)
    assert I.SAE_REGISTRY["b_sae"].kwargs == {"input_dim": 512, "hidden_dim": 32768, "gamma": 1.5, "n_bits": 4}
    assert I.SAE_REGISTRY["q_sae"].kwargs == {"input_dim": 512, "hidden_dim": 32768, "top_k": 32, "abs_range": 1.5,
                                              "n_bits": 4, "allow_bias": True}
    assert I.SAE_REGISTRY["rq_sae"].kwargs["n_bits"] == 4 and I.SAE_REGISTRY["baseline_sae"].kwargs["hidden_dim"] == 32768
    assert I.available_saes(tmp_path)["b_sae"] == tmp_path / "Trained_SAEs" / "b_sae_32768_4_bits.pth"
    with pytest.raises(KeyError):
        I.load_sae("nope")
    with pytest.raises(FileNotFoundError):
        I.load_sae("b_sae", checkpoint_root=tmp_path, device="cpu")
    # checkpoint round trip on a small configuration (kwargs overridden, device cpu: weights only)
    src = Q.BinarySAE(64, 1024, 1.5, 4)
    torch.save(src.state_dict(), tmp_path / "b.pth")
    w = I.load_sae("b_sae", device="cpu", checkpoint_path=tmp_path / "b.pth", input_dim=64, hidden_dim=1024)
    assert isinstance(w.model, Q.BinarySAE) and not w.model.training
    assert torch.equal(w.model.decoder.weight, src.decoder.weight)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        w(torch.zeros(2, 64))
    with pytest.raises(ValueError):
        w([])
    with pytest.raises(TypeError):
        w("x")
    # EleutherAI safetensors layout -> baseline_sae layout (:253-271)
    from safetensors.torch import save_file

    e = {"encoder.weight": torch.randn(128, 16), "encoder.bias": torch.randn(128), "W_dec": torch.randn(128, 16),
         "b_dec": torch.randn(16)}
    save_file(e, str(tmp_path / "sae.safetensors"))
    wb = I.load_sae("baseline_sae", device="cpu", checkpoint_path=tmp_path / "sae.safetensors", input_dim=16, hidden_dim=128)
    assert torch.equal(wb.model.decoder.weight, e["W_dec"].t()) and torch.equal(wb.model.encoder[0].weight, e["encoder.weight"])
    d = wb.decoder_dictionary()
    assert set(d) == {"weight", "bias"} and tuple(d["weight"].shape) == (16, 128)
    assert remap_eleuther_keys({"encoder.0.weight": 1}) == {"encoder.0.weight": 1}
    rq = Q.ResidualQuantizedSAE(16, 64, 32, 1.5, 3)
    torch.save({"state_dict": rq.state_dict()}, tmp_path / "rq.pth")     # wrapped checkpoints are unwrapped
    wr = I.load_sae("rq_sae", device="cpu", checkpoint_path=tmp_path / "rq.pth", input_dim=16, hidden_dim=64, n_bits=3)
    assert set(wr.decoder_dictionary()) >= {"level_0_weight", "level_2_effective_weight", "level_0_bias"}


def test_analysis_has_no_cpu_path():
    """quantizedsae_b200.analysis mirrors scripts/analysis/dynamic_analysis.py but runs on CUDA only."""
    from quantizedsae_b200 import analysis as AN
    from quantizedsae_b200.inference import framework as FW

    for name in ("compute_reconstruction_error", "compute_reconstruction_error_by_level", "compute_l0_by_level",
                 "compute_activation_stats", "analyze_dataset", "_activation_mask", "_hidden_dim"):
        assert callable(getattr(AN, name))
    m = Q.BinarySAE(64, 2048, 4.0, 4)
    sae = FW.SAEWrapper(FW.SAE_REGISTRY["b_sae"], m, "cpu")
    assert AN._hidden_dim(sae) == 2048
    with pytest.raises(RuntimeError, match="CUDA only"):
        AN.compute_reconstruction_error(sae, [torch.zeros(4, 64)], device="cpu")


def test_sharded_transport_defaults_and_k_send():
    from quantizedsae_b200.sharded import DictionaryShardedBinarySAE, ShardPlan

    m = DictionaryShardedBinarySAE(64, 4096, 4.0, 4, rank=0, world_size=1, ops=False)
    assert m.transport == "nccl" and m._peer is None and m.trim_min_k == 256
    p = ShardPlan(2 ** 20, 8, 5)
    # never more than the shard's own top-k, never less than its fair share
    for k in (32, 255, 256, 2097, 4096):
        assert p.k_send(k) <= p.k_local(k) and p.k_send(k) >= min(p.k_local(k), -(-k // 8))


def test_param_key_handles_inference_tensors_and_invalidate_exists():
    """`_version` raises for tensors created under inference_mode; every module offers invalidate() for
    in-place edits made through `.data` (which bump no version counter)."""
    import torch

    import quantizedsae_b200 as Q
    from quantizedsae_b200.sae.base import PreparedCache, param_key

    with torch.inference_mode():
        w = torch.zeros(4)
    assert param_key(w)[0][2] == -1
    v = torch.zeros(4)
    k0 = param_key(v)
    v.add_(1)
    assert param_key(v) != k0                  # ordinary in-place edits are seen ...
    k1 = param_key(v)
    v.data.mul_(2)
    assert param_key(v) == k1                  # ... edits through .data are not: invalidate() is the contract
    for m in (Q.BinarySAE(16, 64, 4.0, 4), Q.BaselineSparseAutoencoder(16, 64), Q.TernarySparseAutoencoder(16, 64),
              Q.QuantizedMatryoshkaSAE(16, 64, 4, n_bits=2), Q.ResidualQuantizedSAE(16, 64, 4, n_bits=2)):
        m._prep = getattr(m, "_prep", PreparedCache())
        m._prep.get("x", (1,), lambda: 1)
        m.invalidate()
        assert m._prep._slots == {}


def test_training_side_has_no_cpu_path_and_validates_arguments():
    """SURVEY 8f-4 entry points: host-side validation of the C ABI and the no-CPU-fallback rule of the module API."""
    import ctypes as C

    lib = _lib.load()
    n = C.c_size_t(0)
    assert lib.qsae_rigl_workspace_bytes(C.byref(n)) == 0 and n.value >= 8192
    assert lib.qsae_rigl_workspace_bytes(None) == -1
    assert lib.qsae_rows_scatter_add(None, None, None, 4, 4, 0, 16, 1.0, None, None, None) == -1          # D <= 0
    assert lib.qsae_rows_scatter_add(None, None, None, 4, 4, 8, 16, 1.0, None, None, None) == -1          # null pointers
    assert lib.qsae_rows_scatter_add(None, None, None, 0, 4, 8, 16, 1.0, None, None, None) == 0           # empty batch
    assert lib.qsae_bsae_logit_grad(None, None, 4, 4, 40, None, 0.0, 0, None, None) == -1                 # n_bits range
    starts = (C.c_int * 3)(0, 5, 9)
    assert lib.qsae_matryoshka_backward_finish(None, None, None, None, None, starts, 2, 10, 4, 0.0, 0, None, None, None) == -1
    assert b"span [0, H]" in lib.qsae_last_error()
    assert lib.qsae_rigl_init_mask(None, None, 70000, 70000, 1, None, 0, None) == -1                      # D * H >= 2^32
    assert b"2^32" in lib.qsae_last_error()
    ste = Q.STEWeights(64, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ste.init_mask(0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ste.update_mask(0.3)
    m = Q.BinarySAE(8, 256, 4.0, 4)
    m.autograd = True
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 8))
    dec = Q.QuantizedMatryoshkaDecoder(64, 8, n_bits=4)
    with pytest.raises(RuntimeError, match="no training context"):
        dec.apply_secant_grad()
