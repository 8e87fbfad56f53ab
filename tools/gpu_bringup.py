"""GPU bring-up script (not a pytest file): exercises every kernel against the numpy oracle and
prints diagnostics. Usage on the GPU box:  python tests/gpu_bringup.py [--big]
"""
import sys
import time
import traceback
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import qsae_oracle as O  # noqa: E402
from quantizedsae_b200 import _lib as L  # noqa: E402
from tests.golden import cases  # noqa: E402

dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
results = {}


def step(name):
    def deco(fn):
        t0 = time.time()
        try:
            fn()
            torch.cuda.synchronize()
            results[name] = "ok"
            print(f"[ok]   {name} ({time.time()-t0:.2f}s)", flush=True)
        except Exception as e:  # noqa: BLE001
            results[name] = f"FAIL {type(e).__name__}: {e}"
            print(f"[FAIL] {name}: {type(e).__name__}: {e}", flush=True)
            traceback.print_exc()
        return fn
    return deco


print("device:", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
lib = L.load()
L.check(lib.qsae_check_device())


@step("cast_bf16")
def _():
    a = np.random.default_rng(0).standard_normal((1000, 77)).astype(np.float32)
    got = L.cast_bf16(T(a)).float().cpu().numpy()
    assert np.array_equal(got, cases.round_bf16(a))


@step("pack_bitplanes + dequant_soft")
def _():
    for name in ["bsae_soft_d64_h2048", "bsae_polar_d64_h2048", "bsae_soft_d32_h1024_8b", "bsae_polar_d256_h8192_2b"]:
        cfg = cases.BSAE_CASES[name]
        inp = cases.bsae_inputs(cfg)
        packed, pol, gap = L.pack_bitplanes(T(inp["logits"]), cfg["D"], cfg["n_bits"])
        iw = O.dequant_hard(inp["logits"], cfg["n_bits"])
        if cfg["n_bits"] <= 4:
            got = O.unpack_nibbles(packed.cpu().numpy())
        else:
            got = packed.cpu().numpy().view(np.int8)
        assert np.array_equal(got, iw), name
        ref_pol = O.polarize_loss(inp["logits"], cfg["n_bits"])
        assert abs(pol - ref_pol) <= 1e-5 * max(1.0, abs(ref_pol)), (pol, ref_pol)
        soft = L.dequant_soft(T(inp["logits"]), cfg["D"], cfg["n_bits"]).cpu().numpy()
        np.testing.assert_allclose(soft, O.dequant_soft(inp["logits"], cfg["n_bits"]), rtol=1e-5, atol=2e-7 * 2 ** cfg["n_bits"])
        print("   ", name, "pol", pol, "gap", gap)


@step("decode int4 / f32 / int8 + densify + transpose")
def _():
    rng = np.random.default_rng(1)
    for (B, k, H, D) in [(37, 4, 2048, 64), (130, 32, 4096, 512), (9, 65, 1024, 256)]:
        iw = rng.integers(-8, 8, size=(H, D)).astype(np.int8)
        vals = rng.standard_normal((B, k)).astype(np.float32)
        idx = np.stack([rng.choice(H, k, replace=False) for _ in range(B)]).astype(np.int32)
        bias = rng.standard_normal(D).astype(np.float32)
        ref = O.decode_rows(vals, idx, iw.astype(np.float32), 0.5, bias)
        got = L.decode_int4(T(vals), T(idx), T(O.pack_nibbles(iw)), H, D, 0.5, T(bias)).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
        got = L.decode_int8(T(vals), T(idx), T(iw), H, D, 0.5, T(bias)).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
        rows = rng.standard_normal((H, D)).astype(np.float32)
        ref = O.decode_rows(vals, idx, rows, 1.0, None)
        got = L.decode_rows_f32(T(vals), T(idx), T(rows), H, D, 1.0, None).cpu().numpy()
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5)
        dense = L.densify(T(vals), T(idx), H).cpu().numpy()
        assert np.array_equal(dense, O.densify(vals, idx, H))
        assert np.array_equal(L.transpose(T(rows)).cpu().numpy(), rows.T)


@step("encode_dense_f32 (SIMT) + topk_dense")
def _():
    rng = np.random.default_rng(2)
    for (B, H, D, k) in [(20, 1000, 72, 5), (33, 4096, 512, 65), (5, 300, 8, 224)]:
        x = rng.standard_normal((B, D)).astype(np.float32)
        W = cases.xavier_uniform(rng, H, D)
        b = (0.01 * rng.standard_normal(H)).astype(np.float32)
        z = L.encode_dense(T(x), T(W), T(b))
        ref = O.encode_pre(x, W, b)
        np.testing.assert_allclose(z.cpu().numpy(), ref, rtol=1e-4, atol=1e-5)
        vals, idx = L.topk_dense(z, k)
        rv, ri = O.topk_rows(z.cpu().numpy(), k)
        assert np.array_equal(idx.cpu().numpy(), ri), (B, H, D, k)
        assert np.array_equal(vals.cpu().numpy(), rv)
    # ties: lowest index first
    z = np.zeros((3, 5000), dtype=np.float32)
    z[:, 2500] = 1
    vals, idx = L.topk_dense(T(z), 4)
    assert idx.cpu().numpy().tolist() == [[2500, 0, 1, 2]] * 3, idx.cpu().numpy().tolist()


def tc_case(B, H, D, k, seed, act=0):
    rng = np.random.default_rng(seed)
    x = cases.round_bf16(rng.standard_normal((B, D)).astype(np.float32))
    W = cases.round_bf16(cases.xavier_uniform(rng, H, D))
    b = (0.01 * rng.standard_normal(H)).astype(np.float32)
    return x, W, b


@step("tcgen05 encoder: dense dump vs fp32 (small)")
def _():
    for (B, H, D) in [(128, 256, 64), (128, 512, 128), (200, 1024, 512), (77, 1000, 72)]:
        x, W, b = tc_case(B, H, D, 1, 3)
        z = L.encode_dense_tc(T(x), L.cast_bf16(T(W)), T(b)).cpu().numpy()
        ref = O.encode_pre(x, W, b)
        err = np.abs(z - ref).max()
        print(f"    B={B} H={H} D={D}: max|err|={err:.3e}  ref rms={ref.std():.3f}", flush=True)
        if not err < 1e-4:
            bad = np.argwhere(np.abs(z - ref) > 1e-4)
            print("    first bad (row, col):", bad[:8].tolist(), "n_bad", len(bad))
            print("    z[0,:8]  ", z[0, :8])
            print("    ref[0,:8]", ref[0, :8])
            raise AssertionError("tensor-core GEMM mismatch")


@step("fused encode_topk vs oracle (small)")
def _():
    for (B, H, D, k) in [(32, 2048, 64, 4), (48, 4096, 512, 8), (300, 8192, 256, 32), (130, 32768, 512, 65)]:
        x, W, b = tc_case(B, H, D, k, 4)
        vals, idx, _ = L.encode_topk(T(x), L.cast_bf16(T(W)), None, T(b), k)
        rv, ri = O.topk_rows(O.encode_pre(x, W, b), k)
        gi = idx.cpu().numpy()
        rows_bad = int((np.sort(gi, 1) != np.sort(ri, 1)).any(1).sum())
        order_bad = int((gi != ri).any(1).sum())
        verr = float(np.abs(vals.cpu().numpy() - rv).max()) if order_bad == 0 else float("nan")
        print(f"    B={B} H={H} D={D} k={k}: rows with wrong set {rows_bad}, wrong order {order_bad}, max|dv| {verr:.2e}", flush=True)
        assert rows_bad == 0 and order_bad == 0
        assert verr < 1e-4


@step("fused encode_topk exact mode (fp32 inputs)")
def _():
    rng = np.random.default_rng(5)
    B, H, D, k = 256, 32768, 512, 32
    x = rng.standard_normal((B, D)).astype(np.float32)
    W = cases.xavier_uniform(rng, H, D)
    b = (0.01 * rng.standard_normal(H)).astype(np.float32)
    vals, idx, flags = L.encode_topk(T(x), L.cast_bf16(T(W)), T(W), T(b), k, exact=True, want_flags=True)
    rv, ri = O.topk_rows(O.encode_pre(x, W, b), k)
    gi = idx.cpu().numpy()
    rows_bad = int((np.sort(gi, 1) != np.sort(ri, 1)).any(1).sum())
    print(f"    exact: rows with wrong set {rows_bad}/{B}, flagged {int(flags.sum())}, max|dv| {np.abs(vals.cpu().numpy()-rv).max():.2e}")
    assert rows_bad == 0
    vals2, idx2, _ = L.encode_topk(T(x), L.cast_bf16(T(W)), None, T(b), k, exact=False)
    rows_bad2 = int((np.sort(idx2.cpu().numpy(), 1) != np.sort(ri, 1)).any(1).sum())
    print(f"    bf16-only on fp32 inputs: rows with wrong set {rows_bad2}/{B} (expected ~14%)")


def bench(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


@step("timing: b_sae 512->32768")
def _():
    H, D = 32768, 512
    for B in ([4096, 65536] if "--big" in sys.argv else [4096]):
        for k in (32, 65):
            g = torch.Generator(device=dev).manual_seed(0)
            x = torch.randn((B, D), device=dev, generator=g).bfloat16().float()
            W = ((torch.rand((H, D), device=dev, generator=g) * 2 - 1) * (6.0 / (H + D)) ** 0.5).bfloat16().float()
            b = torch.zeros(H, device=dev)
            Wb = L.cast_bf16(W)
            packed = torch.randint(0, 256, (H, D // 2), dtype=torch.uint8, device=dev, generator=g)
            bd = torch.randn(D, device=dev, generator=g)
            t_enc = bench(lambda: L.encode_topk(x, Wb, None, b, k))
            vals, idx, _ = L.encode_topk(x, Wb, None, b, k)
            t_dec = bench(lambda: L.decode_int4(vals, idx, packed, H, D, 0.5, bd))
            t_ex = bench(lambda: L.encode_topk(x, Wb, W, b, k, exact=True))
            fl = 2.0 * B * H * D
            print(f"    B={B} k={k}: encode+select {t_enc*1e3:.1f} us ({fl/t_enc/1e9:.1f} TFLOP/s, {B/t_enc/1e3:.2f} Mtok/s)"
                  f" | exact {t_ex*1e3:.1f} us | decode_int4 {t_dec*1e3:.1f} us", flush=True)


print("\nSUMMARY")
for k_, v in results.items():
    print(f"  {k_:55s} {v}")
sys.exit(0 if all(v == "ok" for v in results.values()) else 1)
