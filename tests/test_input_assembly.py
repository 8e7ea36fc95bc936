"""Host-side input assembly (C++, tendermintx_b200/csrc/input.cu) vs the Python restatement in oracle/tm_inputs.py:
byte-identical blobs.  Runs on CPU (no kernel is launched).  Real mocha-4 fixtures are used when /root/reference is
present (authoring container); synthetic chains written to a temp dir run everywhere."""
import json
import os

import pytest

import tendermintx_b200 as tmx
from oracle import tm_inputs as ti

REF_FIX = "/root/reference/circuits/fixtures/mocha-4"
HERE = os.path.dirname(os.path.abspath(__file__))


def test_synthetic_chain_roundtrip(tmp_path):
    for seed, n, n_max, kw in [(0, 5, 8, {}), (1, 16, 16, {"rnd": 3}), (2, 9, 12, {"absent_frac": 0.3}), (3, 130, 256, {})]:
        src, t, g = ti.synthetic_source(seed=seed, n_validators=n, **kw)
        d = tmp_path / f"s{seed}"
        src.write(str(d))
        f = tmx.InputDataFetcher(d)
        th = ti.header_hash(src.signed_header(t)["header"])
        assert f.header_hash(t) == th
        assert f.get_skip_inputs(n_max, t, th, g) == ti.skip_inputs(src, n_max, t, th, g)
    src, t, g = ti.synthetic_source(seed=5, n_validators=7, step=True)
    d = tmp_path / "step"
    src.write(str(d))
    th = ti.header_hash(src.signed_header(t)["header"])
    assert tmx.InputDataFetcher(d).get_step_inputs(8, t, th) == ti.step_inputs(src, 8, t, th)


def test_reference_assertions_become_errors(tmp_path):
    src, t, g = ti.synthetic_source(seed=0, n_validators=5)
    src.write(str(tmp_path))
    f = tmx.InputDataFetcher(tmp_path)
    th = ti.header_hash(src.signed_header(t)["header"])
    with pytest.raises(tmx.TmxError) as e:  # REF input/mod.rs:439-444
        f.get_skip_inputs(4, t, th, g)
    assert e.value.code == 1 and "VALIDATOR_SET_SIZE_MAX" in str(e.value)
    with pytest.raises(tmx.TmxError) as e:  # REF input/mod.rs:450-455
        f.get_skip_inputs(8, t, bytes(32), g)
    assert "sanity check" in str(e.value)
    with pytest.raises(tmx.TmxError) as e:  # REF input/mod.rs:211 (missing fixture)
        f.get_skip_inputs(8, t, th, g + 1)
    assert e.value.code == 4


@pytest.mark.skipif(not os.path.isdir(REF_FIX), reason="reference fixtures only exist in the authoring container")
def test_mocha4_fixtures_match_golden_blobs():
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as fh:
        cases = {c["name"]: c for c in json.load(fh)["cases"]}
    f = tmx.InputDataFetcher(REF_FIX)
    for name, (kind, n_max, a, b) in {"skip_3000_3100_n4": ("skip", 4, 3000, 3100), "skip_10000_10500_n32": ("skip", 32, 10000, 10500),
                                      "step_10500_n4_with_dummy": ("step", 4, 10500, None), "step_10000_n2": ("step", 2, 10000, None)}.items():
        h = f.header_hash(a)
        blob = f.get_skip_inputs(n_max, a, h, b) if kind == "skip" else f.get_step_inputs(n_max, a, h)
        assert blob.hex() == cases[name]["blob"], name
