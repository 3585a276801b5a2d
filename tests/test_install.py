"""CPU, build container only (skipped where /root/reference is absent): the host mirror keeps the
reference's signatures and `install()` rebinds the hot-path symbols of an importable `ultralytics`."""
import inspect
import os
import sys

import pytest

from util import ROOT
import ycr_b200  # noqa: F401

REF = "/root/reference/ultralytics-main"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    return make_golden.import_reference()


def _params(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]


def test_signatures_match_reference(ref):
    rloss, rtal, rops, rhead = ref
    from ycr_b200 import tal, ops, head, loss
    assert _params(ops.non_max_suppression) == _params(rops.non_max_suppression)
    ours = _params(tal.TaskAlignedAssigner.forward)
    assert ours[:-1] == _params(rtal.TaskAlignedAssigner.forward) and ours[-1] == ("grid", None)
    assert _params(tal.TaskAlignedAssigner.__init__) == _params(rtal.TaskAlignedAssigner.__init__)
    assert _params(head.Segment.__init__) == _params(rhead.Segment.__init__)
    assert _params(loss.MaskIOULoss.forward) == _params(rloss.MaskIOULoss.forward)
    assert [n for n, _ in _params(loss.v8SegmentationLoss.__call__)] == \
        [n for n, _ in _params(rloss.v8SegmentationLoss.__call__)]
    assert _params(tal.make_anchors_polar) == _params(rtal.make_anchors_polar)


def test_install_rebinds_symbols(ref):
    rloss, rtal, rops, rhead = ref
    from ycr_b200 import install, tal, ops, loss, head
    saved = {(m, a): getattr(sys.modules[m], a) for m, a, _, _ in install.PATCHES if m in sys.modules
             and hasattr(sys.modules[m], a)}
    try:
        done = install.install()
        assert ("ultralytics.utils.ops", "non_max_suppression") in done
        assert rops.non_max_suppression is ops.non_max_suppression
        assert rtal.TaskAlignedAssigner is tal.TaskAlignedAssigner
        assert rloss.v8SegmentationLoss is loss.v8SegmentationLoss
        assert rhead.Segment is head.Segment
    finally:
        for (m, a), v in saved.items():
            setattr(sys.modules[m], a, v)
