"""CPU: the analysis oracle (oracle/analysis_oracle.py) against the outputs of the unmodified reference
analysis functions (tests/golden/analysis_*.npz, scripts/analysis/dynamic_analysis.py run through the
reference's own SAEWrapper). Counts, masks, co-activation and token lists bit exact; MSE 2e-5 relative
(the reference sums float32 squares per batch)."""
import numpy as np
import pytest

from oracle import analysis_oracle as A
from tests import analysis_common as AC
from tests.golden import cases


@pytest.mark.parametrize("name", list(AC.ANALYSIS_CASES))
def test_analysis_oracle_matches_reference(golden_dir, name):
    kind, cfg = AC.ANALYSIS_CASES[name]
    g = np.load(golden_dir / f"{name}.npz")
    inp = AC.inputs(kind, cfg)
    assert cases.checksum({k: v for k, v in inp.items() if k != "stages"}) == str(g["input_sha"])
    res = A.analyze(kind, cfg, inp, AC.batches(inp["x"]), g["token_ids"], AC.TOKENS_PER_CONTEXT)
    assert np.array_equal(res["mask"], AC.golden_mask(g, cfg["H"]))
    AC.check_against_golden(res, g, cfg["H"], mse_rtol=2e-5)
