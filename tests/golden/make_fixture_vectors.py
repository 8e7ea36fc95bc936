"""Builds tests/golden/fixture_vectors.json from the reference's mocha-4 fixtures (run in the authoring
container, where /root/reference exists): for each case the public input bytes, the packed off-chain blob
(include/tmx_types.h) produced by oracle/tm_inputs.py, and the expected output header, which is the block hash
recorded in the fixture itself (commit.block_id.hash) -- an answer neither implementation computed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tm_inputs as ti  # noqa: E402

FIX = "/root/reference/circuits/fixtures/mocha-4"


def main():
    fx, sb = ti.FixtureSource(FIX), ti.SignedBlockSource(FIX)
    cases = []

    def add(name, kind, src, n_max, a, b=None):
        h = ti.header_hash(src.signed_header(a)["header"])
        if kind == "skip":
            blob = ti.skip_inputs(src, n_max, a, h, b)
            pub = ti.skip_public_input(a, h, b)
            target = b
        else:
            blob = ti.step_inputs(src, n_max, a, h)
            pub = ti.step_public_input(a, h)
            target = a + 1
        expected = src.signed_header(target)["commit"]["block_id"]["hash"].lower()
        cases.append({"name": name, "kind": kind, "n_max": n_max, "chain_id": "mocha-4", "input": pub.hex(),
                      "blob": blob.hex(), "expected_output": expected})

    # REF circuits/skip.rs:190-282, circuits/step.rs:172-268
    add("skip_3000_3100_n4", "skip", fx, 4, 3000, 3100)
    add("skip_10000_10500_n4", "skip", fx, 4, 10000, 10500)
    add("skip_10000_10500_n32", "skip", fx, 32, 10000, 10500)
    add("step_3000_n4", "step", fx, 4, 3000)
    add("step_10000_n2", "step", fx, 2, 10000)
    add("step_10500_n4_with_dummy", "step", fx, 4, 10500)
    add("step_157000_n128", "step", sb, 128, 157000)
    add("skip_15000_50000_n128", "skip", sb, 128, 15000, 50000)
    with open(os.path.join(ROOT, "tests", "golden", "fixture_vectors.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_fixture_vectors.py", "cases": cases}, f, indent=0)
    for c in cases:
        print(c["name"], len(c["blob"]) // 2, c["expected_output"])


if __name__ == "__main__":
    main()
