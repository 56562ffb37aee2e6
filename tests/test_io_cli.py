"""HDF5 subset reader/writer, AMISR-file reader + quality filter, coefficient-file layout, CLI surface.
CPU only (no kernel calls)."""
import io
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden
import ref_port as rp


def test_h5lite_roundtrip(tmp_path):
    from volumetricinterp_b200 import h5lite
    fn = str(tmp_path / "t.h5")
    a = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    i = np.arange(6, dtype=np.int64).reshape(3, 2)
    with h5lite.Writer(fn) as h5:
        h5.array("/UnixTime", i)
        h5.group("/Coeffs", title="Dataset")
        h5.array("/Coeffs/C", a)
        h5.array("/Deep/er/x", np.float32([1.5, 2.5]))
        h5.strings("/FitParams/reglist", ["curvature", "0thorder"])
        h5.string("/FitParams/regmethod", b"chi2")
        h5.string("/ConfigFile/Contents", "[DEFAULT]\nA = 1\n".encode())
        h5.array("/empty", np.zeros((0, 3)))
    with h5lite.File(fn) as h5:
        assert h5.keys("/") == ["Coeffs", "ConfigFile", "Deep", "FitParams", "UnixTime", "empty"]
        assert np.array_equal(h5["/Coeffs/C"], a) and h5["/Coeffs/C"].dtype == np.float64
        assert np.array_equal(h5["/UnixTime"], i) and h5["/UnixTime"].dtype == np.int64
        assert np.array_equal(h5["/Deep/er/x"], np.float32([1.5, 2.5]))
        assert list(h5["/FitParams/reglist"]) == [b"curvature", b"0thorder"]
        assert h5["/FitParams/regmethod"] == b"chi2"
        assert h5["/ConfigFile/Contents"].decode() == "[DEFAULT]\nA = 1\n"
        assert h5["/empty"].shape == (0, 3)
        # PyTables node attributes (what the reference's reader uses to pick the returned Python type)
        at = h5.attrs("/Coeffs/C")
        assert at["CLASS"] == b"ARRAY" and at["FLAVOR"] == b"numpy"
        assert h5.attrs("/FitParams/regmethod")["FLAVOR"] == b"python"
        assert h5.attrs("/Coeffs")["TITLE"] == b"Dataset" and h5.attrs("/")["PYTABLES_FORMAT_VERSION"] == b"2.1"
    raw = open(fn, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    assert int.from_bytes(raw[40:48], "little") == len(raw)      # end-of-file address in the superblock


def _config(tmp_path, fn_in, fn_out, model_extra=""):
    cfg = tmp_path / "config.ini"
    cfg.write_text(f"""[DEFAULT]
PARAM = dens
FILENAME = {fn_in}
OUTPUTFILENAME = {fn_out}
REGULARIZATION_LIST = curvature
REGULARIZATION_METHOD = chi2
ERRLIM = 1e10,1e13
GOODFITCODE = 1,2,3,4
CHI2LIM = 0.1,10

[MODEL]
NAME = sphharmlag
MAXK = 2
MAXL = 2
CAP_LIM = 10
MAX_Z_INT = INF
LATCP = 78
LONCP = 262
{model_extra}
[VALIDATE]
STARTTIME = 2016-11-27T22:46:00
ENDTIME = 2016-11-27T22:49:00
ALTITUDES = 250.0,300.0
COLORLIM = 0.0,5.0e11
OUTPNGNAME = test_fig.png
""")
    return str(cfg)


def test_read_datafile_and_quality_filter_match_oracle(tmp_path):
    """AMISR-layout file -> (utime, lat, lon, alt, value, error): same gate mask as interpolate.py:645-664."""
    from volumetricinterp_b200 import Interpolate, synth
    om = rp.SphHarmLag(2, 2, 10, 78, 262)
    fn = str(tmp_path / "amisr.h5")
    arrays = synth.write_amisr_file(fn, 5, 12, 4, A_of=om.basis, seed=11)
    it = Interpolate(_config(tmp_path, fn, str(tmp_path / "out.h5")))
    utime, lat, lon, alt, value, error = it.read_datafile(fn)
    rl, ro, ra, rv, re_ = rp.quality_filter(arrays["/Geomag/Latitude"], arrays["/Geomag/Longitude"],
                                            arrays["/Geomag/Altitude"], arrays["/FittedParams/Ne"],
                                            arrays["/FittedParams/dNe"], arrays["/FittedParams/FitInfo/chi2"],
                                            arrays["/FittedParams/FitInfo/fitcode"], it.errlim, it.chi2lim, it.goodfitcode)
    assert np.array_equal(utime, arrays["/Time/UnixTime"])
    for a, b in ((lat, rl), (lon, ro), (alt, ra), (value, rv), (error, re_)):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.isnan(value).any() and np.isfinite(value).any()
    assert np.array_equal(np.isnan(value), np.isnan(error))       # identical gate masks


def test_config_keys_and_errors(tmp_path):
    from volumetricinterp_b200 import Interpolate
    it = Interpolate(_config(tmp_path, "in.h5", "out.h5"))
    assert it.regularization_list == ["curvature"] and it.reg_method == "chi2" and it.param == "dens"
    assert it.errlim == [1e10, 1e13] and it.chi2lim == [0.1, 10.0] and it.goodfitcode == [1, 2, 3, 4]
    assert it.model.nbasis == 8 and set(it.model.eval_reg_matricies) == {"curvature", "0thorder"}
    # a regulariser the model does not offer -> KeyError, as interpolate.py:488-493
    it.regularization_list = ["tikhonov"]
    with pytest.raises(KeyError):
        it.eval_reg_matrices()


def test_regularisation_matrices_match_reference():
    """Host-side Omega / Psi (separable quadratures evaluated once each) == the reference's N(N+1)/2 triple quads."""
    import warnings
    from volumetricinterp_b200.models import sphharmlag
    g = load_golden("lo12_two")
    m = sphharmlag.Model(io.StringIO(g["config_text"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert np.array_equal(m.eval_omega(), g["reg_curvature"])
        assert np.array_equal(m.eval_psi(), g["reg_0thorder"])


def test_cli_surface():
    from volumetricinterp_b200 import run_volumetricinterp as cli
    with pytest.raises(SystemExit):
        cli.main(["--help"])
    with pytest.raises(SystemExit):
        cli.main([])          # config_file is required, as in the reference


def _chunked_fixture(path, datasets):
    """A minimal HDF5 file with CHUNKED datasets, assembled byte by byte from the HDF5 file-format specification
    (version 0 superblock, version 1 object headers, data layout message v3 class 2, filter pipeline message v1,
    version 1 B-tree of raw-data chunks: spec sections III.A.1, IV.A.2.i/l, III.A.1 "B-link trees") -- the form
    AMISR fitted files use for their big arrays.  Only the group scaffolding is shared with h5lite.Writer; the
    chunk index, the filter message and the filters themselves (shuffle, deflate, fletcher32) are written here,
    independently of the reader under test.  datasets: {name: (array, chunk dims, [filter ids], two_level)}."""
    import struct
    import zlib
    from volumetricinterp_b200 import h5lite

    def fletcher32(data):
        if len(data) % 2:
            data = data + b"\x00"
        w = np.frombuffer(data, dtype=">u2").astype(np.uint64)
        s1 = s2 = 0
        for x in w:                              # reference definition (HDF5 H5checksum.c), fine for tiny chunks
            s1 = (s1 + int(x)) % 65535
            s2 = (s2 + s1) % 65535
        return struct.pack("<I", (s2 << 16) | s1)

    class W(h5lite.Writer):
        def _emit_data(self, node):
            if "chunks" not in node:
                return super()._emit_data(node)
            a, cd, filt, two = node["arr"], node["chunks"], node["filters"], node["two_level"]
            rank, es = a.ndim, a.dtype.itemsize
            grid = [range(0, a.shape[d], cd[d]) for d in range(rank)]
            import itertools
            entries = []
            for offs in itertools.product(*grid):
                chunk = np.zeros(cd, dtype=a.dtype)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cd, a.shape))
                chunk[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
                raw = chunk.tobytes()
                for fid in filt:
                    if fid == 2:                                     # shuffle: byte planes
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, es).T.tobytes()
                    elif fid == 1:
                        raw = zlib.compress(raw, 6)
                    elif fid == 3:
                        raw = raw + fletcher32(raw)
                entries.append((offs, len(raw), self._alloc(raw)))

            def node_bytes(level, items):                            # items: (offsets, size, child address)
                out = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), h5lite.UNDEF, h5lite.UNDEF))
                for offs, size, child in items:
                    out += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
                    out += struct.pack("<Q", child)
                last = tuple(a.shape[d] + cd[d] for d in range(rank))    # final key: beyond every chunk
                out += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", o) for o in last) + struct.pack("<Q", 0)
                return bytes(out)

            if two and len(entries) > 2:
                half = len(entries) // 2
                kids = [entries[:half], entries[half:]]
                leaves = [(k[0][0], 0, self._alloc(node_bytes(0, k))) for k in kids]
                btree = self._alloc(node_bytes(1, leaves))
            else:
                btree = self._alloc(node_bytes(0, entries))
            names = {1: b"deflate", 2: b"shuffle", 3: b"fletcher32"}
            fm = struct.pack("<BB6x", 1, len(filt))
            for fid in filt:
                nm = names[fid] + b"\x00"
                nm = nm + b"\x00" * (-len(nm) % 8)
                cdv = {1: [6], 2: [es], 3: []}[fid]
                fm += struct.pack("<HHHH", fid, len(nm), 0, len(cdv)) + nm + b"".join(struct.pack("<I", v) for v in cdv)
                if len(cdv) % 2:
                    fm += b"\x00" * 4
            layout = struct.pack("<BBBQ", 3, 2, rank + 1, btree) + b"".join(struct.pack("<I", c) for c in cd) + \
                struct.pack("<I", es)
            msgs = [h5lite._msg(0x0001, h5lite._space_msg(a.shape)), h5lite._msg(0x0003, h5lite._dtype_msg(a.dtype), flags=1),
                    h5lite._msg(0x0005, struct.pack("<BBBB", 2, 1, 0, 0)), h5lite._msg(0x000B, fm), h5lite._msg(0x0008, layout)]
            return self._alloc(self._header(msgs))

    with W(path, pytables_attrs=False) as h5:
        for name, (arr, cd, filt, two) in datasets.items():
            h5.array(name, arr)
            parent, leaf = h5._parent(name)
            parent["members"][leaf].update(chunks=tuple(cd), filters=list(filt), two_level=two)


def test_h5lite_reads_chunked_filtered_datasets(tmp_path):
    """Reader paths the writer never produces: chunked layout through a v1 chunk B-tree (one and two levels, edge
    chunks that overhang the dataset), shuffle + deflate + fletcher32 filter pipeline, 3-D float64 and 2-D int64."""
    from volumetricinterp_b200 import h5lite
    rng = np.random.default_rng(0)
    ne = rng.standard_normal((7, 5, 9)) * 1e11
    ne[2, 3, :4] = np.nan
    code = rng.integers(0, 6, (7, 11)).astype(np.int64)
    alt = np.linspace(1e5, 7e5, 55).reshape(5, 11)
    fn = str(tmp_path / "chunked.h5")
    _chunked_fixture(fn, {"/FittedParams/Ne": (ne, (3, 2, 4), [2, 1, 3], True),
                          "/FittedParams/FitInfo/fitcode": (code, (4, 4), [2, 1], False),
                          "/Geomag/Altitude": (alt, (5, 11), [], False),
                          "/Time/UnixTime": (np.arange(14, dtype=np.float64).reshape(7, 2), (7, 2), [1], False)})
    with h5lite.File(fn) as h5:
        assert np.array_equal(h5["/FittedParams/Ne"], ne, equal_nan=True)
        assert np.array_equal(h5["/FittedParams/FitInfo/fitcode"], code) and h5["/FittedParams/FitInfo/fitcode"].dtype == np.int64
        assert np.array_equal(h5["/Geomag/Altitude"], alt)
        assert np.array_equal(h5["/Time/UnixTime"], np.arange(14, dtype=np.float64).reshape(7, 2))
