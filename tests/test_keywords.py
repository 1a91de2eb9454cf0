"""Reset-time keyword factories against the reference's notebook-printed goldens."""
import numpy as np
import pytest

from adcraft_b200 import keywords as kwm


def _rng(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def test_manual_bidding_example_notebook_keywords():
    """manual_bidding_example.ipynb:84-87: mean_volume 16, cvr 0.5, 2 keywords, reset(seed=0)."""
    t = kwm.sample_implicit_keywords_from_quantiles(2, _rng(0), {"mean_volume": 16, "conversion_rate": 0.5})
    assert t.describe() == (
        "kw0 params:\n volume: (16, 1),   imp_intercept: 0.6459721981904619,   imp_slope: 9.492169932038324,"
        "   bctr: 0.7526828432972257,   sctr: 0.5,   mean revenue: 1.229655446429944,   std revenue: 0.3184237989333203\n"
        "kw1 params:\n volume: (16, 8),   imp_intercept: 0.7565469048855986,   imp_slope: 4.711917056603076,"
        "   bctr: 0.10219080013611848,   sctr: 0.5,   mean revenue: 0.5459178688435826,   std revenue: 0.08867800662594895")


def test_example_compute_metrics_notebook_keywords():
    """example_compute_metrics.ipynb:57,70: mean_volume 100, cvr 0.3, N=30, reset(seed=10)."""
    t = kwm.sample_implicit_keywords_from_quantiles(30, _rng(10), {"mean_volume": 100, "conversion_rate": 0.3})
    rows = [
        ((100, 13), 0.7386049044669925, 7.3627261809468685, 0.31804044252579394, 0.3, 0.9703987841419266, 0.10969622240554196),
        ((100, 31), 0.8124585400585577, 4.2044508665737235, 0.11163888030443188, 0.3, 1.0705702906229062, 0.03917634894645788),
        ((100, 8), 0.3221557501686997, 10.852394187688585, 0.25535694693155775, 0.3, 0.48794031199293697, 0.10905562626755198),
        ((100, 13), 0.6633304743341284, 12.371355808003809, 0.7229296935496455, 0.3, 0.6057705058105929, 0.13625750505976256),
        ((100, 33), 0.42493114913951424, 20.570283754642272, 0.24720905165502582, 0.3, 0.7614144153214522, 0.05960807585712174),
        ((100, 39), 0.5136281208098616, 14.963304547637959, 0.2649683118527094, 0.3, 1.0178599609968022, 0.1859930520278975),
        ((100, 8), 0.8114621475355906, 30.02868444366422, 0.8255108524903837, 0.3, 0.5383433924839526, 0.15456012353158874),
    ]
    for k, (vol, loc, inv_scale, bctr, sctr, rev, rev_sd) in enumerate(rows):
        assert (t.vol_mean[k], t.vol_std[k]) == vol
        assert t.p1[k] == loc and 1.0 / t.p2[k] == inv_scale
        assert t.ctr[k] == bctr and t.cvr[k] == sctr and t.rev_mean[k] == rev and t.rev_std[k] == rev_sd


def test_random_explicit_keywords_ranges():
    """gymnasium_kw_utils.py:129-131 precedence quirk: mean volume in [14, 29]."""
    t = kwm.sample_random_keywords(500, _rng(3))
    assert t.kind == kwm.EXPLICIT
    assert t.vol_mean.min() >= 14 and t.vol_mean.max() <= 29
    assert np.all(t.vol_std <= 0.5 * (t.vol_mean + 1))
    assert np.all((t.p2 >= 0) & (t.p2 <= 25)) and np.all((t.p1 >= 0) & (t.p1 <= 1.5))


def test_per_env_keyword_sets_and_table_helpers():
    t = kwm.sample_implicit_keywords_from_quantiles(5, _rng(1), {"mean_volume": 64, "conversion_rate": 0.1},
                                                    num_envs=3)
    assert t.per_env and t.vol_mean.shape == (3, 5) and t.env_stride == 5
    assert not np.array_equal(t.p1[0], t.p1[1])
    assert t.env(2).vol_mean.shape == (5,)
    shared = kwm.sample_implicit_keywords_from_quantiles(5, _rng(1), {"mean_volume": 64, "conversion_rate": 0.1})
    assert np.array_equal(shared.p1, t.p1[0])  # env 0 of the per-env draw is the shared draw


def test_quantile_table_sources():
    cols = kwm.quantile_rows_from_config({"mean_volume": 128, "conversion_rate": 0.8})
    assert cols["min_vol"][0] == cols["max_vol"][0] == 128 and cols["median_sctr"][0] == 0.8
    assert cols["min_ave_cpc"][0] == 0.3 and cols["max_std_rpsc"][0] == 0.3
    custom = kwm.quantile_rows_from_config({"quantile_table": {k: v for k, v in cols.items()}})
    a = kwm.sample_implicit_keywords_from_quantiles(4, _rng(2), {"mean_volume": 128, "conversion_rate": 0.8})
    b = kwm.sample_implicit_keywords_from_quantiles(4, _rng(2), {"quantile_table": custom})
    for n in kwm.PARAM_NAMES:
        assert np.array_equal(getattr(a, n), getattr(b, n))


def test_quantile_csv_wire_format_round_trip(tmp_path):
    """count_/min_/median_/max_<param> columns behind an unnamed index column (what pandas' to_csv
    writes, experiment_quantiles.py:68-73): written, read back, and used through the reference's
    three loading conventions (make/load_quant_func pair, quantiles_folder default loader)."""
    rng = np.random.default_rng(0)
    rows = 4
    cols = {}
    for p in kwm.QUANTILE_PARAMS:
        lo = np.round(rng.uniform(0.05, 0.4, rows), 3)  # short decimals: every float parser agrees on them
        cols[f"count_{p}"] = np.array([3.0, 0.0, 5.0, 2.0]) if p == "bctr" else np.full(rows, 3.0)
        cols[f"min_{p}"], cols[f"median_{p}"], cols[f"max_{p}"] = lo, lo + 0.125, lo + 0.5
    cols["min_vol"], cols["median_vol"], cols["max_vol"] = (np.array([8.0, 20, np.nan, 90]),
                                                            np.array([12.0, 40, np.nan, 120]),
                                                            np.array([16.0, 64, np.nan, 256]))
    folder = tmp_path / "q" / "setA"
    folder.mkdir(parents=True)
    kwm.write_quantile_csv(cols, str(folder / "auction_data.csv"))
    first = open(folder / "auction_data.csv").readline().strip().split(",")
    assert first[0] == "" and first[1:5] == ["count_vol", "min_vol", "median_vol", "max_vol"]
    back = kwm.read_quantile_csv(str(folder / "auction_data.csv"))
    assert set(back) == set(cols)
    for c in cols:
        np.testing.assert_array_equal(back[c], cols[c])
    cfg = {"outer_directory": str(tmp_path / "q") + "/", "quantiles_folder": "setA/"}
    a = kwm.sample_implicit_keywords_from_quantiles(12, _rng(3), cfg)                       # default CSV loader
    b = kwm.sample_implicit_keywords_from_quantiles(12, _rng(3), {"quantile_table": cols})  # in memory
    for n in kwm.PARAM_NAMES:
        np.testing.assert_array_equal(getattr(a, n), getattr(b, n))
    assert np.any(a.vol_mean == 0)  # the NaN-volume bucket gives zero-volume keywords (utils:296-300)
    with pytest.raises(AssertionError, match="Invalid quantile parameters"):
        kwm.sample_implicit_keywords_from_quantiles(2, _rng(3), {"outer_directory": str(tmp_path) + "/",
                                                                 "quantiles_folder": "missing/"})
    # experiment configs: make_quant_func writes <outer>/<vol>_<cvr>.csv, load_quant_func reads it
    ecfg = {"outer_directory": str(tmp_path), "mean_volume": 64, "conversion_rate": 0.1,
            "make_quant_func": kwm.make_experiment_quantiles, "load_quant_func": kwm.load_experiment_quantiles}
    c = kwm.sample_implicit_keywords_from_quantiles(7, _rng(4), ecfg)
    d = kwm.sample_implicit_keywords_from_quantiles(7, _rng(4), {"mean_volume": 64, "conversion_rate": 0.1})
    assert (tmp_path / "64_0.1.csv").exists()
    for n in kwm.PARAM_NAMES:
        np.testing.assert_array_equal(getattr(c, n), getattr(d, n))
