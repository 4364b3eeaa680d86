"""The launch-group plan of the host-buffer pipeline (orbx_plan_groups: the same next_group() orbx_extract_batch and
orbx_extract_batch_multi use, device-free).  Round 1 had a ramp rule that could make a group larger than max_batch -- larger
than the workspace and the staging slots -- for n_frames just above max_batch or just above the ramp sums; this sweeps every
n_frames around those values."""
import ctypes as C

import numpy as np
import pytest


def plan(n_frames, max_batch, consumers=1, ramp=1):
    import extractorb_b200 as ex
    L = ex.load_library()
    L.orbx_plan_groups.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    buf = np.zeros(6000, np.int32)
    n = L.orbx_plan_groups(n_frames, max_batch, consumers, ramp, buf.ctypes.data, len(buf))
    assert 0 <= n <= len(buf)
    return buf[:n].tolist()


@pytest.mark.parametrize("max_batch", [1, 4, 31, 32, 33, 64, 100, 256, 512])
def test_groups_cover_all_frames_and_never_exceed_max_batch(max_batch):
    for n_frames in list(range(0, 3 * max_batch + 40)) + [480 + k for k in range(0, 40)] + [992 + k for k in range(0, 40)] + [4096, 4097, 5000]:
        for consumers in (1, 2, 3, 8):
            g = plan(n_frames, max_batch, consumers)
            assert sum(g) == n_frames, (n_frames, max_batch, consumers, g)
            assert all(0 < x <= max_batch for x in g), (n_frames, max_batch, consumers, g)


def test_ramp_shape():
    """Host pipelines start with small groups (the first H2D has nothing to overlap with) and end with small ones."""
    g = plan(4096, 256)
    assert g[:4] == [32, 64, 128, 256] and g[-1] <= 128 and max(g) == 256
    assert plan(4096, 256, ramp=0) == [256] * 16
    assert plan(40, 256) == [32, 8] or sum(plan(40, 256)) == 40
    assert plan(20, 16) == [16, 4]                     # groups of <= 32 frames are not ramped
    assert plan(0, 16) == []


def test_bad_arguments():
    import extractorb_b200 as ex
    L = ex.load_library()
    L.orbx_plan_groups.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    assert L.orbx_plan_groups(10, 0, 1, 1, None, 0) == -2
    assert L.orbx_plan_groups(-1, 4, 1, 1, None, 0) == -2
    assert L.orbx_plan_groups(10, 4, 0, 1, None, 0) == -2
    assert L.orbx_plan_groups(10, 4, 1, 1, None, 0) == 3
