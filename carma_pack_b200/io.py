"""Light-curve ingestion: ASCII `time y yerr` files (the format of cpp_tests/data/*.dat,
examples/OGLE-LMC-LPV-00007.dat and src/paper/data/*.txt in the reference) and FITS binary tables (Kepler / RXTE
light curves, src/paper/data/kepler_zw229_Q7.fits read at src/paper/carma_paper.py:522-531) -> cleaned arrays ->
ragged CSR packing for MultiSeries.  numpy only: the FITS reader below parses the standard directly
(astropy is not a dependency)."""
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


_TFORM = {"L": "i1", "B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "E": ">f4", "D": ">f8"}


def _fits_header(buf, pos):
    """Parse one header (80-byte cards in 2880-byte blocks) starting at byte `pos`: (dict, offset of the data unit)."""
    hdr = {}
    while True:
        block = buf[pos:pos + 2880]
        if len(block) < 2880:
            raise ValueError("truncated FITS header")
        pos += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                return hdr, pos
            if card[8:10] != "= ":
                continue
            val = card[10:]
            if val.lstrip().startswith("'"):
                v = val.lstrip()[1:]
                v = v[:v.find("'")].rstrip() if "'" in v else v.rstrip()
            else:
                v = val.split("/")[0].strip()
                if v in ("T", "F"):
                    v = (v == "T")
                else:
                    try:
                        v = int(v)
                    except ValueError:
                        try:
                            v = float(v.replace("D", "E"))
                        except ValueError:
                            pass
            hdr[key] = v


def read_fits_table(path, hdu=1):
    """Columns of the binary-table extension number `hdu` (1 = first extension) of a FITS file, as a dict of numpy
    arrays in native byte order.  Supports the scalar and fixed-repeat column formats L, B, I, J, K, E, D and A."""
    buf = open(path, "rb").read()
    pos, index = 0, 0
    while pos < len(buf):
        hdr, data_pos = _fits_header(buf, pos)
        naxis = int(hdr.get("NAXIS", 0))
        nbytes = 0
        if naxis > 0:
            nbytes = abs(int(hdr["BITPIX"])) // 8
            for k in range(1, naxis + 1):
                nbytes *= int(hdr["NAXIS%d" % k])
            nbytes = (nbytes + int(hdr.get("PCOUNT", 0))) * int(hdr.get("GCOUNT", 1))
        if index == hdu:
            if hdr.get("XTENSION") != "BINTABLE":
                raise ValueError("HDU %d of %s is not a binary table" % (hdu, path))
            fields = []
            for k in range(1, int(hdr["TFIELDS"]) + 1):
                form = str(hdr["TFORM%d" % k]).strip()
                rep = "".join(ch for ch in form if ch.isdigit())
                code = form[len(rep)]
                rep = int(rep) if rep else 1
                name = str(hdr.get("TTYPE%d" % k, "COL%d" % k)).strip()
                if code == "A":
                    fields.append((name, "S%d" % rep))
                elif code in _TFORM:
                    fields.append((name, _TFORM[code]) if rep == 1 else (name, _TFORM[code], (rep,)))
                else:
                    raise ValueError("unsupported TFORM %r" % form)
            dt = np.dtype(fields)
            nrow = int(hdr["NAXIS2"])
            if dt.itemsize != int(hdr["NAXIS1"]):
                raise ValueError("row size mismatch: %d != NAXIS1 %d" % (dt.itemsize, int(hdr["NAXIS1"])))
            tab = np.frombuffer(buf, dtype=dt, count=nrow, offset=data_pos)
            return {n: np.ascontiguousarray(tab[n]).astype(tab[n].dtype.newbyteorder("=")) for n in dt.names}
        pos = data_pos + ((nbytes + 2879) // 2880) * 2880
        index += 1
    raise ValueError("%s has no HDU %d" % (path, hdu))


def read_fits(path, time_col="TIME", flux_col="SAP_FLUX", err_col="SAP_FLUX_ERR", hdu=1, zero_time=True):
    """A light curve from a FITS binary table (column names are case-insensitive), cleaned as in
    src/paper/carma_paper.py:522-531: rows with non-finite time / flux dropped, time measured from its minimum."""
    cols = {k.upper(): v for k, v in read_fits_table(path, hdu).items()}
    t, y, e = (np.asarray(cols[c.upper()], dtype=float) for c in (time_col, flux_col, err_col))
    t, y, e = clean_light_curve(t, y, e)
    if zero_time and t.size:
        t = t - t.min()
    return t, y, e


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
