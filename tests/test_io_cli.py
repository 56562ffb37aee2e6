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
