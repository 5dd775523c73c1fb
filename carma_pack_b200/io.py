"""Light-curve ingestion: ASCII `time y yerr` files (the format of cpp_tests/data/*.dat,
examples/OGLE-LMC-LPV-00007.dat and src/paper/data/*.txt in the reference) -> cleaned arrays ->
ragged CSR packing for MultiSeries."""
import numpy as np


def clean_light_curve(time, y, yerr):
    """Cleaning of src/paper/carma_paper.py:23-43: drop non-finite rows, sort by time, keep the first
    of any duplicated time (CarmaModel.__init__, carma_pack.py:32-35)."""
    time, y, yerr = (np.asarray(a, dtype=float) for a in (time, y, yerr))
    ok = np.isfinite(time) & np.isfinite(y) & np.isfinite(yerr)
    time, y, yerr = time[ok], y[ok], yerr[ok]
    order = np.argsort(time, kind="stable")
    time, y, yerr = time[order], y[order], yerr[order]
    keep = np.concatenate([[True], np.diff(time) > 0])
    return time[keep], y[keep], yerr[keep]


def read_ascii(path, usecols=(0, 1, 2), **kw):
    """Read a whitespace-separated light-curve file and clean it."""
    data = np.loadtxt(path, usecols=usecols, **kw)
    return clean_light_curve(data[:, 0], data[:, 1], data[:, 2])


def pack_ragged(curves, min_points=2):
    """Concatenate (time, y, yerr) triples into the CSR layout of carma_multi_series_create.
    Returns (time, y, yerr, offsets, kept_indices); curves shorter than min_points are skipped."""
    ts, ys, es, off, kept = [], [], [], [0], []
    for i, (t, y, e) in enumerate(curves):
        t, y, e = clean_light_curve(t, y, e)
        if t.size < min_points:
            continue
        ts.append(t); ys.append(y); es.append(e)
        off.append(off[-1] + t.size)
        kept.append(i)
    if not ts:
        raise ValueError("no usable light curves")
    return (np.concatenate(ts), np.concatenate(ys), np.concatenate(es), np.asarray(off, dtype=np.int64),
            np.asarray(kept, dtype=np.int64))


def load_multi_series(paths, device=0, **kw):
    """Read many ASCII light curves into one device-resident ragged batch."""
    from ._lib import MultiSeries
    t, y, e, off, kept = pack_ragged([read_ascii(p, **kw) for p in paths])
    return MultiSeries(t, y, e, off, device=device), kept
