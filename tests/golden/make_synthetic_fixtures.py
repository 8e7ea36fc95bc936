"""Writes the deterministic synthetic chains used by bench.py and the full-size GPU tests (SURVEY.md section 8d:
no mocha-4 fixture has more than 100 validators and the fixtures' chain id fails CelestiaConfig's check), in the
reference's RPC-JSON fixture layout: tests/golden/celestia/<case>/<height>/{commit.json, validators_<page>.json}.
Generator: oracle/tm_inputs.py synthetic_source (keys = SHA-256("tmx/val" || seed || i), libsodium signatures)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tm_inputs as ti  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "celestia")
CASES = [(f"skip_n128_seed{s}", dict(seed=s, n_validators=128)) for s in range(8)]
CASES += [("step_n128_seed0", dict(seed=0, n_validators=128, step=True)), ("skip_n256_seed0", dict(seed=0, n_validators=256)),
          ("skip_n16_seed0", dict(seed=0, n_validators=16))]

if __name__ == "__main__":
    shutil.rmtree(OUT, ignore_errors=True)
    index = {}
    for name, kw in CASES:
        src, t, g = ti.synthetic_source(**kw)
        src.write(os.path.join(OUT, name))
        th = ti.header_hash(src.signed_header(t)["header"])
        index[name] = {"trusted": t, "target": g, "trusted_hash": th.hex(),
                       "target_hash": ti.header_hash(src.signed_header(g)["header"]).hex(), "n_validators": kw["n_validators"],
                       "kind": "step" if kw.get("step") else "skip"}
    import json
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(index, f, indent=1)
    print({k: v["target_hash"][:12] for k, v in index.items()})
