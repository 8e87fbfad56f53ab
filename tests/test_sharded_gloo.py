"""N > 1 host logic of the dictionary-sharded b_sae on CPU: world_size-2 gloo processes run the real
collective choreography (all-gather of candidates, deterministic merge, owner-only decode,
reduce-scatter of partial reconstructions) with the numpy ops backend, and must reproduce the
single-process oracle forward over the full dictionary."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import qsae_oracle as O
from quantizedsae_b200.sharded import DictionaryShardedBinarySAE, ShardPlan
from tests.sharded_common import NumpyShardOps, full_state_dict, sharded_case


def test_shard_plan_arithmetic():
    p = ShardPlan(2 ** 20, 8, 3)
    assert p.shard_latents == 131072 and p.latent_range() == (393216, 524288) and p.latent_begin == 393216
    assert p.k_local(32) == 32 and p.k_local(2097) == 2097 and ShardPlan(64, 8, 0).k_local(32) == 8
    assert p.k_send(32) == 32 and p.k_send(2097) == 376 and ShardPlan(2 ** 17, 8, 0).k_send(262) == 84
    assert ShardPlan(2 ** 20, 1, 0).k_send(2097) == 2097 and ShardPlan(4096, 8, 0).k_send(4000) == 512   # never above k_local
    assert p.padded_batch(4096) == 4096 and p.padded_batch(37) == 40
    assert [ShardPlan(64, 4, r).row_range(10) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert ShardPlan(64, 4, 3).row_range(5) == (5, 5)            # trailing ranks may own no row
    with pytest.raises(ValueError):
        ShardPlan(100, 8, 0)
    with pytest.raises(ValueError):
        ShardPlan(64, 4, 4)
    cfg, inp = sharded_case()
    sd = ShardPlan(cfg["H"], 4, 2).shard_state_dict(full_state_dict(inp), cfg["n_bits"])
    assert tuple(sd["encoder.0.weight"].shape) == (512, 64) and tuple(sd["decoder.weight"].shape) == (512, 256)
    assert torch.equal(sd["encoder.0.bias"], torch.from_numpy(inp["be"][1024:1536]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(B, skew):
    """skew: every latent of the first half of the dictionary gets a large encoder bias, so shard 0 owns ALL
    winners -- the truncated candidate exchange must notice that its list was used up and fall back."""
    cfg, inp = sharded_case(B=B)
    if skew:
        inp["be"] = inp["be"].copy()
        inp["be"][: cfg["H"] // 2] += np.float32(8.0)
    return cfg, inp


def _worker(rank, world, port, k_frac, B, gather, skew, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg, inp = _case(B, skew)
        m = DictionaryShardedBinarySAE(cfg["D"], cfg["H"], cfg["gamma"], cfg["n_bits"], ops=False)
        m.ops = NumpyShardOps(m)
        m.load_state_dict(m.plan.shard_state_dict(full_state_dict(inp), cfg["n_bits"]), strict=True)
        m.k = k_frac
        m.gather_output = gather
        with torch.no_grad():
            lat, rows, pol = m(torch.from_numpy(inp["x"]))
        out[rank] = (lat.values.numpy(), lat.indices.numpy(), rows.numpy(), float(pol), m.plan.row_range(B),
                     m.last_exchange)
    finally:
        dist.destroy_process_group()


# k = 4 / 32 (full exchange); k = 409 of 2048 latents: truncated exchange (307 candidates per shard); the same with
# a skewed dictionary: shard 0 owns every winner, the merge reports its list used up, full exchange follows
@pytest.mark.parametrize("world,k_frac,B,gather,skew,exchange", [
    (2, 0.002, 37, False, False, "full"), (2, 2 ** -6, 16, False, False, "full"), (2, 0.002, 9, True, False, "full"),
    (2, 0.2, 11, False, False, "truncated"), (2, 0.2, 7, True, True, "full")])
def test_dictionary_sharded_forward_under_gloo(world, k_frac, B, gather, skew, exchange):
    import torch.multiprocessing as mp

    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, k_frac, B, gather, skew, out), nprocs=world, join=True)
        res = {r: out[r] for r in range(world)}
    cfg, inp = _case(B, skew)
    k = int(cfg["H"] * k_frac)
    assert all(res[r][5] == exchange for r in range(world))
    rv, ri, rr, rp = O.bsae_forward(inp["x"], inp["We"], inp["be"], inp["logits"], inp["bd"], n_bits=cfg["n_bits"],
                                    gamma=cfg["gamma"], k=k, mode="hard")
    for r in range(world):
        v, i, rows, pol, (a, b), _ = res[r]
        assert np.array_equal(i, ri), f"rank {r}: global top-k indices differ from the single-process oracle"
        np.testing.assert_array_equal(v, rv)
        ref_rows = rr if gather else rr[a:b]
        assert rows.shape == ref_rows.shape
        rms = float(np.sqrt(np.mean(rr.astype(np.float64) ** 2)))
        np.testing.assert_allclose(rows, ref_rows, rtol=1e-4, atol=1e-4 * rms)
        assert pol == pytest.approx(rp, abs=1e-12)
    assert [res[r][4] for r in range(world)] == [ShardPlan(cfg["H"], world, r).row_range(B) for r in range(world)]
