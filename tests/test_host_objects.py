"""
CPU tests of the host-side mirror of the reference's objects.  The cases follow
the reference's own unit tests (code/tests/test_*.py: range validation, profile
bounds, root-pdf normalisation, closure shapes / round trip) and add checks
against the golden outputs of the reference's closures.
"""
import os

import numpy as np
import pytest

from hydromodel_b200 import pack
from hydromodel_b200.src.soil_properties import SoilProperties
from hydromodel_b200.src.water_content import WaterContent
from hydromodel_b200.src.hydraulic_conductivity import HydraulicConductivity
from hydromodel_b200.src.porosity import Porosity
from hydromodel_b200.src.tree_roots import TreeRoots
from hydromodel_b200.src.utilities import logN_rnd, find_wtd, lerp_table
from hydromodel_b200.src.models.vrettas_fung import VrettasFung
from hydromodel_b200.src.models.vanGenuchten import vanGenuchten
from hydromodel_b200.src.models.hydrological_model import HydrologicalModel
from helpers import column_from_json, rhs_cases, rhs_configs


# ---- value objects (reference tests: test_soil_properties / water_content / hydraulic_conductivity)
def test_soil_properties_validation():
    s = SoilProperties()
    assert s.m == pytest.approx(0.5)
    for attr, bad in (("n", 0.0), ("alpha", 0.0), ("psi_sat", 1.0), ("epsilon", 0.0)):
        with pytest.raises(ValueError):
            setattr(s, attr, bad)
    with pytest.raises(ValueError):
        SoilProperties(n=0.0)
    assert SoilProperties(psi_sat=5.0).psi_sat == 0.0          # constructor clamps
    assert SoilProperties(epsilon=0.0).epsilon == 1.0e-8


def test_water_content_validation():
    w = WaterContent()
    with pytest.raises(ValueError):
        w.max += 1.0
    with pytest.raises(ValueError):
        w.min -= 1.0
    with pytest.raises(ValueError):
        w.res = w.min
    assert w._checkValues()
    assert w.mid == pytest.approx(0.5 * (w.min + w.max))
    with pytest.raises(ValueError):
        WaterContent(0.3, 0.1, 0.05)
    w.wlt, w.flc = -1000.0, 300.0
    assert (w.wlt, w.flc) == (-1000.0, 300.0)


def test_hydraulic_conductivity_validation():
    k = HydraulicConductivity()
    for attr in ("sat_soil", "sat_saprolite", "sat_fresh_bedrock"):
        with pytest.raises(ValueError):
            setattr(k, attr, 0.0)
        with pytest.raises(ValueError):
            HydraulicConductivity(**{attr: -1.0})
    for attr in ("sigma_noise", "lambda_exponent"):
        with pytest.raises(ValueError):
            setattr(k, attr, -0.1)
    assert HydraulicConductivity(sigma_noise=-3.0).sigma_noise == 0.0


# ---- utilities (reference test_utilities.py)
def test_logn_rnd_mean_and_shapes():
    rng = np.random.default_rng(0)
    x = logN_rnd(20.5, 1.0, rng.standard_normal(25000))
    assert np.mean(x) == pytest.approx(20.5, rel=1e-3)
    assert logN_rnd(np.ones(7), np.ones(7), np.zeros(7)).shape == (7,)
    assert logN_rnd(np.ones((5, 3)), np.ones((5, 3)), np.zeros((5, 3))).shape == (5, 3)


def test_find_wtd_known_answers():
    assert find_wtd(np.r_[np.zeros(9, bool), True]) == 9
    assert find_wtd(np.ones(10, bool)) == 0
    assert find_wtd(np.array([0, 0, 1, 0, 1, 0, 1, 1, 1, 1], bool)) == 6
    assert find_wtd(np.array([1, 1, 1, 0, 1, 1, 0, 1, 1, 1], bool)) == 7
    assert find_wtd(None) is None


def test_lerp_table_matches_scipy_interp1d():
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(1)
    xk = np.cumsum(rng.uniform(0.5, 2.0, 40)); yk = rng.standard_normal((3, 40))
    xq = rng.uniform(xk[0], xk[-1], 200)
    assert np.array_equal(lerp_table(xk, yk, xq), interp1d(xk, yk)(xq))
    with pytest.raises(ValueError):
        lerp_table(xk, yk[0], xk[-1] + 1.0)


# ---- porosity / roots (reference test_porosity.py, test_tree_roots.py)
Z = np.arange(0.0, 2005.0, 5.0)
LAYERS = (0.0, 50.0, 200.0, 2000.0)


@pytest.mark.parametrize("profile", ["Constant", "Linear", "Exponential", "Stratified", "Noisy"])
def test_porosity_profiles(profile):
    theta, soil = WaterContent(0.05, 0.30, 0.04, 1500.0, 340.0), SoilProperties()
    p = Porosity(Z, LAYERS, theta, soil, profile)
    full = p()
    assert len(full) == 3 and all(v.shape == Z.shape for v in full)
    assert np.all(p.profile >= theta.min) and np.all(p.profile <= theta.max)
    assert np.asarray(p(12.5)).shape == (3,)
    assert np.asarray(p([12.5, 100.0, 1999.0])).shape == (3, 3)
    assert np.all(p.wilting_point <= p.field_cap)
    assert p.layers == LAYERS


def test_porosity_errors():
    theta, soil = WaterContent(), SoilProperties()
    with pytest.raises(ValueError):
        Porosity(Z, LAYERS, theta, soil, "Nope")
    with pytest.raises(ValueError):
        Porosity(np.array([]), LAYERS, theta, soil, "Constant")
    with pytest.raises(RuntimeError):
        Porosity(Z[::-1], LAYERS, theta, soil, "Constant")


@pytest.mark.parametrize("model", ["Uniform", "Negative_Exp", "Gamma_pdf", "Mixture"])
def test_tree_roots_pdf_and_efficiency(model):
    theta, soil = WaterContent(0.05, 0.30, 0.04, 1500.0, 340.0), SoilProperties()
    por = Porosity(Z, LAYERS, theta, soil, "Stratified")
    roots = TreeRoots(200, 5.0, model, por)
    assert np.sum(roots()) * 5.0 == pytest.approx(1.0, rel=1e-5)
    zr = np.arange(0.0, 1000.0, 5.0)
    rho, water = roots.efficiency(np.zeros(zr.size), zr)
    assert water == 0.0 and np.all(rho == 0.0)
    rng = np.random.default_rng(2)
    th = por(zr + 1e-9)[2] + rng.uniform(0.01, 0.2, zr.size)
    rho, water = roots.efficiency(th, zr + 1e-9)
    if np.any(rho > 0):
        assert np.sum(rho) * 5.0 == pytest.approx(1.0, rel=1e-5)
    with pytest.raises(RuntimeError):
        roots.efficiency(np.zeros(3), zr)
    with pytest.raises(ValueError):
        TreeRoots(200, 5.0, "Nope", por)


def test_gamma_pdf_matches_scipy():
    from scipy.stats import gamma
    from hydromodel_b200.src.tree_roots import _gamma_pdf
    x = np.linspace(1, 100, 50)
    assert np.allclose(_gamma_pdf(x, 2.5, 5.0), gamma.pdf(x, a=2.5, scale=5.0), rtol=1e-13)


# ---- closures on the host (reference test_hydrological_models.py + golden outputs)
def test_host_closures_against_reference(golden):
    g = golden("rhs_cases.npz")
    n = 0
    for cfg in rhs_configs(g):
        col = column_from_json(g[f"{cfg}/params_json"])
        o = col["objs"]
        cls = VrettasFung if o["model_id"] == 0 else vanGenuchten
        hm = cls(o["soil"], o["porosity"], o["K"], o["theta"].res, o["dz"])
        for _, psi, nr, precip, wtd, spin, p in list(rhs_cases(g, cfg))[:6]:
            q, K, C, kb, qinf = hm(psi, col["x"], {"n_rnd": nr})
            ref = g[p + "nodes"]
            for mine, r in zip((q, K, C, kb), ref):
                assert np.allclose(mine, r, rtol=1e-11, atol=0)
            assert float(qinf) == pytest.approx(float(g[p + "qinf"]), rel=1e-11)
            n += 1
    assert n >= 40


def test_pressure_head_round_trip_and_shapes():
    theta, k = WaterContent(0.05, 0.30, 0.04, 1500.0, 340.0), HydraulicConductivity()
    soil = SoilProperties(psi_sat=-0.0047)
    por = Porosity(Z, LAYERS, theta, soil, "Stratified")
    for cls in (VrettasFung, vanGenuchten):
        hm = cls(soil, por, k, theta.res, 5.0)
        th0 = 0.8 * por.profile
        psi, s_eff = hm.pressure_head(th0, Z)
        assert psi.shape == Z.shape and s_eff.shape == Z.shape
        q, K, C, kb, qinf = hm(psi, Z, {"n_rnd": np.zeros(Z.size)})
        assert abs(np.mean(q - th0)) < 0.1                      # the reference's (loose) bar
        assert np.allclose(q, th0, rtol=1e-9)
        psi2 = np.repeat(psi[:, None], 4, axis=1)
        q2, K2, C2, kb2, _ = hm(psi2, Z, {"n_rnd": np.zeros(Z.size)})
        assert q2.shape == (Z.size, 4)
        with pytest.raises(ValueError):
            hm(psi[:-1], Z, {"n_rnd": np.zeros(Z.size)})
    base = HydrologicalModel(soil, por, k, theta.res, 5.0)
    sat_psi, _ = base.pressure_head(por.profile, Z)
    assert np.array_equal(sat_psi, np.arange(Z.size) * 5.0)        # saturated column: spin-up start


def test_pack_quirks():
    """Q1/Q2 bookkeeping of the static pack."""
    from helpers import default_column
    col = default_column()
    par, stat = col["par"], col["stat"]
    assert par[pack.P_PHI_TOP] == col["objs"]["porosity"].profile[0]
    assert par[pack.P_DSMIN_ALL] == (col["objs"]["porosity"].profile - col["objs"]["theta"].res).min()
    assert set(np.unique(stat[pack.S_COEF_MID, :-1])) == {0.05, 0.10, 1.0}
    assert int(par[pack.P_FLAGS]) == pack.FLAG_LF | pack.FLAG_ET
    # midpoint 0 (z = 2.5 cm) lies in the soil layer: coefficient 0.05, mean Ksat_soil
    assert stat[pack.S_COEF_MID, 0] == 0.05 and stat[pack.S_KMEAN_MID, 0] == 8.5


# ----------------------------------------------------------------------------- HDF5 files
def _h5_sample():
    rng = np.random.default_rng(5)
    return {"psi_press": rng.standard_normal((14, 401)), "status": rng.integers(0, 3, 365).astype(np.int8),
            "counters": rng.integers(0, 1 << 40, (4, 9)), "empty": np.zeros((0, 5)),
            "scalar": np.float64(3.5), "f32": rng.standard_normal((3, 4, 5)).astype(np.float32),
            "u16": (np.arange(70000) % 65536).astype(np.uint16), "flag": np.array([True, False, True]),
            "wtd_est_cm": np.round(rng.uniform(500, 1100, (6, 41)))}


def test_hdf5_writer_reader_round_trip(tmp_path):
    """Result files without h5py: every dataset comes back with its
    shape, dtype and bits; gzip level 4 like h5py's compression='gzip'."""
    import struct
    import zlib
    from hydromodel_b200.src import hdf5_min
    data = _h5_sample()
    for kwargs in ({}, {"compression_level": None}, {"chunk_rows": 3}):
        path = str(tmp_path / "out.h5")
        hdf5_min.write_file(path, data, **kwargs)
        back = hdf5_min.read_file(path)
        assert sorted(back) == sorted(data)
        for k, v in data.items():
            v = np.asarray(v)
            want = v.astype(np.uint8) if v.dtype == np.bool_ else v
            assert back[k].shape == want.shape and back[k].dtype == want.dtype, (k, kwargs)
            assert np.array_equal(back[k], want), (k, kwargs)
        buf = open(path, "rb").read()
        # superblock version 0: signature, 8-byte offsets and lengths, end-of-file address
        assert buf[:8] == b"\x89HDF\r\n\x1a\n" and buf[8] == 0 and buf[13] == 8 and buf[14] == 8
        assert struct.unpack_from("<Q", buf, 40)[0] == len(buf) and len(buf) % 8 == 0
        root_oh, = struct.unpack_from("<Q", buf, 64)
        assert buf[root_oh] == 1 and struct.unpack_from("<H", buf, root_oh + 16)[0] == 0x0011
    # the chunk of a compressed dataset is a plain zlib stream of the C-ordered array
    path = str(tmp_path / "one.h5")
    arr = np.arange(12.0).reshape(3, 4)
    hdf5_min.write_file(path, {"a": arr})
    buf = open(path, "rb").read()
    comp = zlib.compress(arr.tobytes(), 4)
    assert buf.count(comp) == 1 and buf.find(comp) % 8 == 0


def test_hdf5_reader_rejects_what_it_does_not_understand(tmp_path):
    from hydromodel_b200.src import hdf5_min
    p = tmp_path / "bad.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(hdf5_min.HDF5FormatError):
        hdf5_min.read_file(str(p))
    good = str(tmp_path / "good.h5")
    hdf5_min.write_file(good, {"a": np.arange(100.0)})
    buf = bytearray(open(good, "rb").read())
    (tmp_path / "trunc.h5").write_bytes(bytes(buf[:len(buf) - 64]))
    with pytest.raises(hdf5_min.HDF5FormatError):
        hdf5_min.read_file(str(tmp_path / "trunc.h5"))
    buf[8] = 3                                             # superblock version 3
    (tmp_path / "v3.h5").write_bytes(bytes(buf))
    with pytest.raises(hdf5_min.HDF5FormatError):
        hdf5_min.read_file(str(tmp_path / "v3.h5"))
    with pytest.raises(hdf5_min.HDF5FormatError):
        hdf5_min.write_file(str(tmp_path / "x.h5"), {"s": np.array(["text"])})


def test_save_and_load_results_like_the_reference(tmp_path):
    """Simulation.saveResults / loadResults (simulation.py:674-746) on top of it."""
    from hydromodel_b200.src.hdf5_io import load_datasets, save_datasets
    data = {k: v for k, v in _h5_sample().items() if k != "flag"}
    path = save_datasets(str(tmp_path / "Sim_00.h5"), data)
    assert path.endswith("Sim_00.h5")
    back = load_datasets(path)
    assert all(np.array_equal(back[k], np.asarray(v)) for k, v in data.items())


def test_hdf5_reader_on_a_file_written_by_libhdf5():
    """The only file written by the real HDF5 library in this image: a MATLAB 7.3 file of scipy's
    test data (512-byte user block, superblock 0, symbol-table root group, version-1 object
    headers, version-1/2 data layout) holding testdouble = pi/4 * arange(9).  The reader walks
    the same structures the writer emits, so this pins the format against libhdf5 itself."""
    import scipy.io
    from hydromodel_b200.src import hdf5_min
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's HDF5 test file is not installed")
    data = hdf5_min.read_file(path)
    assert list(data) == ["testdouble"]
    assert data["testdouble"].dtype == np.float64 and data["testdouble"].shape == (9, 1)
    assert np.array_equal(data["testdouble"].ravel(), np.pi / 4 * np.arange(9.0))


def test_hdf5_against_h5py_when_available(tmp_path):
    """round trips through h5py in both directions -- skipped where h5py is absent
    (this image); runs wherever the reference's own dependency is installed."""
    h5py = pytest.importorskip("h5py")
    from hydromodel_b200.src import hdf5_min
    data = {k: np.asarray(v) for k, v in _h5_sample().items() if k not in ("flag", "scalar")}
    ours = str(tmp_path / "ours.h5")
    hdf5_min.write_file(ours, data)
    with h5py.File(ours, "r") as fh:                       # libhdf5 reads what we write
        assert sorted(fh) == sorted(data)
        for k, v in data.items():
            assert fh[k].shape == v.shape and fh[k].dtype == v.dtype and np.array_equal(fh[k][...], v), k
            if v.size:
                assert fh[k].compression == "gzip"
    theirs = str(tmp_path / "theirs.h5")
    with h5py.File(theirs, "w") as fh:                     # we read what the reference writes
        for k, v in data.items():
            fh.create_dataset(k, data=v, shape=v.shape, compression="gzip")
    back = hdf5_min.read_file(theirs)
    assert all(np.array_equal(back[k], v) for k, v in data.items())


def test_hdf5_size_check_before_a_run():
    from hydromodel_b200.src import hdf5_min
    hdf5_min.check_writable([(65536, 11, 401), (0, 5), ()])
    with pytest.raises(hdf5_min.HDF5FormatError):
        hdf5_min.check_writable([(65536, 3650, 401)])      # 767 GB: beyond 64 chunks of < 4 GiB
