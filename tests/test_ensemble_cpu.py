"""
CPU tests of the ensemble host layer: the vectorised packer must equal the
per-column packer bit for bit, members depend on their global id only, a list of
parameter dicts is regrouped into struct-of-arrays form, and the command-line
entry point keeps the reference's exit-code contract where no GPU is needed.
"""
import copy
import json
import time

import numpy as np
import pytest

from hydromodel_b200 import pack
from hydromodel_b200.sharding import column_range
from hydromodel_b200.synthetic import (default_params, ensemble_columns, ensemble_member,
                                       synthetic_forcing)
from conftest import SITE


@pytest.mark.parametrize("kw", [dict(), dict(perturb_theta=True), dict(mixed=True, perturb_theta=True)])
def test_pack_ensemble_equals_pack_column_bit_for_bit(kw):
    table = synthetic_forcing(120)
    base = default_params(SITE)
    b, cols = ensemble_columns(base, 500, 48, **kw)
    ens = pack.pack_ensemble(b, table, cols, block=20)        # several blocks
    assert ens["n_slow"] == 0
    for k in range(48):
        p, mult = ensemble_member(base, 500 + k, **kw)
        one = pack.pack_column(p, table, precip_mult=mult)
        assert np.array_equal(one["par"], ens["par"][k]), k
        assert np.array_equal(one["stat"], ens["stat"][k]), k
    assert np.array_equal(ens["x"], one["x"])
    for key in ("precipitation_cm", "wtd_idx", "daylight", "atm", "zWtd_cm"):
        assert np.array_equal(ens["forcing"][key], one["forcing"][key])


def test_pack_ensemble_slow_path_for_invalid_and_complex_columns():
    """Columns the array path cannot reproduce go through pack_column: an invalid
    soil group (the reference falls back to defaults, simulation.py:146-160) and the
    default wilting point -1500 with a non-even n (complex field capacity, quirk Q7)."""
    table = synthetic_forcing(30)
    base = default_params(SITE)
    cols = {"n": np.array([2.0, 0.9, 2.2, 2.0]), "a0": np.array([0.009, 0.009, 0.009, -1.0])}
    with np.errstate(all="ignore"):
        ens = pack.pack_ensemble(base, table, cols)
    assert ens["n_slow"] == 3
    for k in range(4):
        p, _ = pack.member_params(base, cols, k)
        with np.errstate(all="ignore"):
            one = pack.pack_column(p, table)
        assert np.array_equal(one["par"], ens["par"][k], equal_nan=True), k
        assert np.array_equal(one["stat"], ens["stat"][k], equal_nan=True), k


def test_split_params_list_regroups_dicts():
    base = default_params(SITE)
    plist = [ensemble_member(base, c, mixed=True, perturb_theta=True)[0] for c in range(9)]
    other = copy.deepcopy(plist[3])
    other["Trees"]["Max_Root_Depth_cm"] = 600.0              # not a per-column key: its own group
    plist.append(other)
    groups = pack.split_params_list(plist)
    assert sorted(len(g[2]) for g in groups) == [1, 9]
    table = synthetic_forcing(20)
    for gbase, cols, idx in groups:
        ens = pack.pack_ensemble(gbase, table, cols)
        for j, i in enumerate(idx):
            one = pack.pack_column(plist[i], table)
            assert np.array_equal(one["par"], ens["par"][j]) and np.array_equal(one["stat"], ens["stat"][j])


def test_members_depend_on_global_id_only_and_shares_tile():
    base = default_params(SITE)
    _, whole = ensemble_columns(base, 0, 101, perturb_theta=True)
    for world in (2, 3, 8):
        parts = [ensemble_columns(base, *column_range(101, r, world), perturb_theta=True)[1] for r in range(world)]
        for key in whole:
            assert np.array_equal(np.concatenate([p[key] for p in parts]), whole[key]), (world, key)
    p, mult = ensemble_member(base, 77, 1024, perturb_theta=True)
    assert p["Soil_Properties"]["n"] == whole["n"][77] and mult == whole["precip_mult"][77]
    assert p["Water_Content"]["Theta_Max"] == whole["Theta_Max"][77]


def test_large_ensemble_packs_in_seconds():
    """65,536 columns must be constructible in < 60 s; 8,192 here (< 8 s)."""
    table = synthetic_forcing(3650)
    t0 = time.perf_counter()
    b, cols = ensemble_columns(default_params(SITE), 0, 8192, perturb_theta=True)
    ens = pack.pack_ensemble(b, table, cols)
    dt = time.perf_counter() - t0
    assert ens["par"].shape == (8192, pack.NPARAM) and ens["stat"].shape == (8192, pack.NSTATIC, 401)
    assert dt < 8.0, dt


def test_cli_exit_code_contract_without_gpu(tmp_path):
    """berkeley_hydro_main.main ends with exit code 1 on every failure
    (code/berkeley_hydro_main.py:65-146): no parameter file, a missing key, a missing CSV."""
    from hydromodel_b200 import berkeley_hydro_main as cli
    with pytest.raises(SystemExit) as exc:
        cli.main(None, None)
    assert exc.value.code == 1
    params = default_params(SITE)
    broken = dict(params)
    del broken["Hydraulic_Conductivity"]
    bad = tmp_path / "bad.json"
    bad.write_text(json.dumps(broken))
    with pytest.raises(SystemExit) as exc:
        cli.main(str(bad), None)
    assert exc.value.code == 1
    good = tmp_path / "good.json"
    good.write_text(json.dumps(params))
    with pytest.raises(SystemExit) as exc:
        cli.main(str(good), str(tmp_path / "missing.csv"))
    assert exc.value.code == 1
    assert cli.validateInputParametersFile(str(good))["Well_No"] == 10
