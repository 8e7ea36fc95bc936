"""the operator's search for the block to request [REF circuits/input/mod.rs:158-186,
circuits/input/tendermint_utils.rs:444-482].  C++ host code behind the C ABI vs the Python restatement, on synthetic
chains with controlled validator-set overlap (run anywhere) and on the reference's mocha-4 fixtures (authoring
container; results also committed as tests/golden/operator_vectors.json)."""
import copy
import json
import os

import pytest

import tendermintx_b200 as tmx
from oracle import tm_inputs as ti

REF_FIX = "/root/reference/circuits/fixtures/mocha-4"
HERE = os.path.dirname(os.path.abspath(__file__))


def _chain_with_overlap(tmp_path, keep, nil_votes=0):
    """start set = the first `keep` validators of the target set plus strangers; returns (dir, start, target)."""
    src, t, g = ti.synthetic_source(seed=11, n_validators=12, absent_frac=0.25)
    target_vals = src.validators(g)
    strangers, _, _ = ti.synthetic_source(seed=12, n_validators=12)
    start_vals = copy.deepcopy(target_vals[:keep]) + copy.deepcopy(strangers.validators(g)[: 12 - keep])
    src.valsets[t] = start_vals
    if nil_votes:
        sigs = src.headers[g]["commit"]["signatures"]
        done = 0
        for s, v in zip(sigs, target_vals):
            if int(s["block_id_flag"]) == 1 and done < nil_votes:  # absent -> nil: carries an address, no block id
                s["block_id_flag"], s["validator_address"] = 3, v["address"]
                done += 1
    d = tmp_path / f"keep{keep}_{nil_votes}"
    src.write(str(d))
    return src, str(d), t, g


@pytest.mark.parametrize("keep,nil_votes", [(0, 0), (1, 0), (2, 0), (3, 0), (5, 0), (12, 0), (2, 3), (1, 3)])
def test_is_valid_skip_matches_oracle_on_synthetic_overlap(tmp_path, keep, nil_votes):
    src, d, t, g = _chain_with_overlap(tmp_path, keep, nil_votes)
    f = tmx.InputDataFetcher(d)
    want = ti.is_valid_skip(src, t, g)
    assert f.is_valid_skip(t, g) == want
    if keep == 0:
        assert not want


def test_full_overlap_all_signing_is_a_valid_skip(tmp_path):
    src, t, g = ti.synthetic_source(seed=2, n_validators=9)
    src.write(str(tmp_path))
    assert ti.is_valid_skip(src, t, g) and tmx.InputDataFetcher(tmp_path).is_valid_skip(t, g)


def test_find_block_to_request_bisects(tmp_path):
    """Only the pairs the reference would touch need fixtures: start, max_end and the midpoints it falls back to."""
    src, t, g = ti.synthetic_source(seed=3, n_validators=8, trusted_height=1000, target_height=1016)
    other, _, _ = ti.synthetic_source(seed=4, n_validators=8, trusted_height=1000, target_height=1016)
    # 1016 and 1008 are signed by strangers (no overlap), 1004 by the start set: the search must return 1004
    hdr, vals = src.headers[1016], src.valsets[1016]
    src.headers[1004], src.valsets[1004] = hdr, vals
    for h in (1016, 1008):
        src.headers[h], src.valsets[h] = other.headers[1016], other.valsets[1016]
    d = tmp_path / "bisect"
    src.write(str(d))
    f = tmx.InputDataFetcher(d)
    assert ti.find_block_to_request(src, 1000, 1016) == 1004
    assert f.find_block_to_request(1000, 1016) == 1004
    assert f.find_block_to_request(1000, 1001) == 1001  # distance one: request a step
    with pytest.raises(tmx.TmxError) as e:  # the reference `expect`s on a missing fixture
        f.find_block_to_request(1000, 1032)
    assert e.value.code == 4
    with pytest.raises(tmx.TmxError):
        f.find_block_to_request(1000, 1000)


@pytest.mark.skipif(not os.path.isdir(REF_FIX), reason="reference fixtures only exist in the authoring container")
def test_reference_fixtures_match_golden_results():
    with open(os.path.join(HERE, "golden", "operator_vectors.json")) as fh:
        gold = json.load(fh)
    f = tmx.InputDataFetcher(REF_FIX)
    src = ti.FixtureSource(REF_FIX)
    for a, b, want in gold["is_valid_skip"]:
        assert f.is_valid_skip(a, b) == want == ti.is_valid_skip(src, a, b), (a, b)
    for a, b, want in gold["find_block_to_request"]:
        assert f.find_block_to_request(a, b) == want == ti.find_block_to_request(src, a, b), (a, b)
    assert any(not w for _, _, w in gold["is_valid_skip"]) and any(w for _, _, w in gold["is_valid_skip"])
