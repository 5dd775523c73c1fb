import numpy as np

from carma_pack_b200 import io as cio


def test_clean_and_pack(tmp_path):
    t = np.array([3.0, 1.0, 2.0, 2.0, np.nan, 5.0])
    y = np.array([30.0, 10.0, 20.0, 21.0, 0.0, np.inf])
    e = np.ones(6)
    tc, yc, ec = cio.clean_light_curve(t, y, e)
    assert list(tc) == [1.0, 2.0, 3.0] and list(yc) == [10.0, 20.0, 30.0]  # first of the duplicated time kept
    p = tmp_path / "lc.dat"
    np.savetxt(p, np.column_stack([t[:4], y[:4], e[:4]]), fmt="%10.5f")
    tr, yr, er = cio.read_ascii(str(p))
    assert list(tr) == [1.0, 2.0, 3.0]
    tt, yy, ee, off, kept = cio.pack_ragged([(t, y, e), (np.array([1.0]), np.array([1.0]), np.array([1.0])),
                                             (np.arange(4.0), np.arange(4.0), np.ones(4))])
    assert list(off) == [0, 3, 7] and list(kept) == [0, 2] and tt.size == 7
    assert np.all(np.diff(tt[:3]) > 0) and np.all(np.diff(tt[3:]) > 0)
