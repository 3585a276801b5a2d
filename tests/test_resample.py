"""Contour resampling: oracle vs the reference's own output (CPU), kernel vs both (GPU)."""
import numpy as np
import pytest
import torch

from util import load_golden
from oracle import polar_oracle as po


def _polys(g):
    out, s = [], 0
    for m in g["sizes"]:
        out.append(g["points"][s:s + int(m)])
        s += int(m)
    return out


def test_oracle_resample_matches_reference():
    g = load_golden("resample")
    res = po.resample_segments(_polys(g), 360)
    assert np.array_equal(np.stack(res), g["out"])
    for r in res:                                   # closed polygon: last point == first point
        assert np.array_equal(r[0], r[-1])


@pytest.mark.gpu
def test_kernel_resample_bit_exact():
    from ycr_b200.ops import resample_segments
    g = load_golden("resample")
    polys = _polys(g)
    res = resample_segments(polys, n=360, device="cuda:0")
    assert len(res) == len(polys)
    assert np.array_equal(torch.stack(res).cpu().numpy(), g["out"])
    # another n, ragged list incl. a 3-point polygon, against the oracle
    rng = np.random.default_rng(11)
    more = [rng.uniform(-5, 700, size=(int(m), 2)).astype(np.float32) for m in (3, 5, 64, 257)]
    for n in (36, 100, 1000):
        got = torch.stack(resample_segments(more, n=n, device="cuda:0")).cpu().numpy()
        assert np.array_equal(got, np.stack(po.resample_segments(more, n)))
    assert resample_segments([], n=360) == []
