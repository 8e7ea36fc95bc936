"""Input assembly, restated in stdlib Python: fixture / RPC-JSON loading, header leaf encoding, header Merkle
proofs, CanonicalVote sign-bytes, per-validator records and the packed off-chain blob (include/tmx_types.h).

TEST INFRASTRUCTURE ONLY.  Follows circuits/input/{mod,conversion,tendermint_utils,utils}.rs of the reference
(cited per function).  Also holds the deterministic synthetic-chain generator of SURVEY.md section 8d, which
emits the same RPC JSON shapes so one loader serves real and synthetic data.
"""
import base64
import hashlib
import json
import os
import struct
from datetime import datetime, timezone

DUMMY_PUBLIC_KEY = bytes.fromhex("3b6a27bcceb6a42d62a3a8d02a6f0d73653215771de243a63ac048a18b59da29")
DUMMY_SIGNATURE = bytes.fromhex(
    "3da1ebdfa96edd181dbe3659d1c051c431f056a5ad6a97a60d5cca10460438783546461e31285fc59f91c7072642745061e2451d5ff33bccd8c3c74dabcaf60a")

KIND_STEP, KIND_SKIP = 0, 1
MAGIC = 0x31584D54
MSG_MAX = 124


def sha256(b):
    return hashlib.sha256(b).digest()


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


# ---- Merkle (REF circuits/input/tendermint_utils.rs:294-372) ----
def leaf_hash(x):
    return sha256(b"\x00" + x)


def inner_hash(l, r):
    return sha256(b"\x01" + l + r)


def split_point(n):
    k = 1
    while k * 2 < n:
        k *= 2
    return k


def merkle_root(items):
    if len(items) == 0:
        return sha256(b"")
    if len(items) == 1:
        return leaf_hash(items[0])
    k = split_point(len(items))
    return inner_hash(merkle_root(items[:k]), merkle_root(items[k:]))


def merkle_proofs(items):
    """Root and, per leaf, the list of aunts from the leaf upwards (REF tendermint_utils.rs:294-336)."""
    def rec(sub):
        if len(sub) == 1:
            return leaf_hash(sub[0]), [[]]
        k = split_point(len(sub))
        lr, lp = rec(sub[:k])
        rr, rp = rec(sub[k:])
        return inner_hash(lr, rr), [p + [rr] for p in lp] + [p + [lr] for p in rp]
    return rec(items)


# ---- header encoding (REF circuits/input/tendermint_utils.rs:374-393) ----
def parse_time(s):
    """RFC 3339 with up to nanosecond precision -> (seconds, nanos)."""
    assert s.endswith("Z")
    body = s[:-1]
    frac = "0"
    if "." in body:
        body, frac = body.split(".")
    secs = int(datetime.strptime(body, "%Y-%m-%dT%H:%M:%S").replace(tzinfo=timezone.utc).timestamp())
    nanos = int((frac + "000000000")[:9])
    return secs, nanos


def enc_timestamp(secs, nanos):
    out = b""
    if secs:
        out += b"\x08" + varint(secs)
    if nanos:
        out += b"\x10" + varint(nanos)
    return out


def enc_bytes_field1(b):
    return (b"\x0a" + varint(len(b)) + b) if b else b""


def enc_block_id(bid):
    """tendermint_proto BlockId: 0a 20 <hash> 12 24 08 <total> 12 20 <hash>."""
    h = bytes.fromhex(bid["hash"])
    parts = bid["parts"]
    psh = b""
    if int(parts["total"]):
        psh += b"\x08" + varint(int(parts["total"]))
    ph = bytes.fromhex(parts["hash"])
    if ph:
        psh += b"\x12" + varint(len(ph)) + ph
    out = b""
    if h:
        out += b"\x0a" + varint(len(h)) + h
    out += b"\x12" + varint(len(psh)) + psh
    return out


def header_leaves(h):
    secs, nanos = parse_time(h["time"])
    ver = b"\x08" + varint(int(h["version"]["block"]))
    if int(h["version"].get("app", 0)):
        ver += b"\x10" + varint(int(h["version"]["app"]))
    hx = lambda k: enc_bytes_field1(bytes.fromhex(h[k]))
    return [
        ver,
        enc_bytes_field1(h["chain_id"].encode()),
        b"\x08" + varint(int(h["height"])),
        enc_timestamp(secs, nanos),
        enc_block_id(h["last_block_id"]),
        hx("last_commit_hash"), hx("data_hash"), hx("validators_hash"), hx("next_validators_hash"),
        hx("consensus_hash"), hx("app_hash"), hx("last_results_hash"), hx("evidence_hash"),
        enc_bytes_field1(bytes.fromhex(h["proposer_address"])),
    ]


def header_hash(h):
    return merkle_root(header_leaves(h))


# ---- validators (REF circuits/input/conversion.rs, tendermint::validator::Info::hash_bytes) ----
def validator_pubkey(v):
    return base64.b64decode(v["pub_key"]["value"])


def validator_bytes(v):
    out = b"\x0a\x22\x0a\x20" + validator_pubkey(v)
    p = int(v["voting_power"])
    if p:
        out += b"\x10" + varint(p)
    return out


def sort_validators(vals):
    """tendermint-rs validator::Set::new ordering: voting power descending, then address ascending."""
    return sorted(vals, key=lambda v: (-int(v["voting_power"]), bytes.fromhex(v["address"])))


def validators_hash(vals):
    return merkle_root([validator_bytes(v) for v in sort_validators(vals)])


# ---- sign bytes (tendermint SignedVote::sign_bytes as used at REF conversion.rs:34-39) ----
def sign_bytes(chain_id, height, rnd, block_id, timestamp):
    body = b"\x08\x02" + b"\x11" + struct.pack("<q", height)
    if rnd:
        body += b"\x19" + struct.pack("<q", rnd)
    if block_id is not None:
        h = bytes.fromhex(block_id["hash"])
        ph = bytes.fromhex(block_id["parts"]["hash"])
        psh = b"\x08" + varint(int(block_id["parts"]["total"])) + b"\x12" + varint(len(ph)) + ph
        cb = b"\x0a" + varint(len(h)) + h + b"\x12" + varint(len(psh)) + psh
        body += b"\x22" + varint(len(cb)) + cb
    ts = enc_timestamp(*parse_time(timestamp))
    body += b"\x2a" + varint(len(ts)) + ts
    body += b"\x32" + varint(len(chain_id)) + chain_id.encode()
    return varint(len(body)) + body


# ---- data sources ----
class FixtureSource:
    """Fixture mode of InputDataFetcher (REF circuits/input/mod.rs:188-282): <root>/<height>/commit.json and
    validators_<page>.json, 100 validators per page."""

    def __init__(self, root):
        self.root = root

    def signed_header(self, height):
        with open(os.path.join(self.root, str(height), "commit.json")) as f:
            return json.load(f)["result"]["signed_header"]

    def validators(self, height):
        vals, page = [], 1
        while True:
            with open(os.path.join(self.root, str(height), f"validators_{page}.json")) as f:
                r = json.load(f)["result"]
            vals.extend(r["validators"])
            if len(vals) >= int(r["total"]):
                return vals
            page += 1


class MemorySource:
    def __init__(self, headers, valsets):
        self.headers, self.valsets = headers, valsets

    def signed_header(self, height):
        return self.headers[height]

    def validators(self, height):
        return self.valsets[height]

    def write(self, root):
        """Emit the same files FixtureSource reads (100 validators per page)."""
        for h, sh in self.headers.items():
            d = os.path.join(root, str(h))
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "commit.json"), "w") as f:
                json.dump({"jsonrpc": "2.0", "id": -1, "result": {"signed_header": sh, "canonical": True}}, f)
            vals = self.valsets[h]
            for p in range(0, max(1, (len(vals) + 99) // 100)):
                page = vals[100 * p:100 * p + 100]
                with open(os.path.join(d, f"validators_{p + 1}.json"), "w") as f:
                    json.dump({"jsonrpc": "2.0", "id": -1, "result": {
                        "block_height": str(h), "validators": page, "count": str(len(page)), "total": str(len(vals))}}, f)


# ---- per-validator records (REF circuits/input/conversion.rs:59-178) ----
def _validator_record(pubkey, sig, msg, msg_len, power, byte_len, signed):
    return struct.pack("<32s32s32s124sIQIB3x", pubkey, sig[:32], sig[32:], msg.ljust(MSG_MAX, b"\x00"), msg_len,
                       power, byte_len, 1 if signed else 0)


def validator_data_from_block(vals, signed_header, n_max):
    hdr, commit = signed_header["header"], signed_header["commit"]
    by_addr = {v["address"]: v for v in vals}
    recs = []
    for i, cs in enumerate(commit["signatures"]):
        v = by_addr.get(vals[i]["address"])
        if v is None:
            continue
        pk, power, blen = validator_pubkey(v), int(v["voting_power"]), len(validator_bytes(v))
        if cs["block_id_flag"] == 2:
            msg = sign_bytes(hdr["chain_id"], int(commit["height"]), int(commit["round"]), commit["block_id"], cs["timestamp"])
            sig = base64.b64decode(cs["signature"])
            assert len(msg) <= MSG_MAX
            recs.append(_validator_record(pk, sig, msg, len(msg), power, blen, True))
        else:
            recs.append(_validator_record(pk, DUMMY_SIGNATURE, b"", 32, power, blen, False))
    while len(recs) < n_max:
        recs.append(_validator_record(DUMMY_PUBLIC_KEY, DUMMY_SIGNATURE, b"", 32, 0, 46, False))
    return recs


def hash_fields_from_block(vals, commit, n_max):
    ordered = sort_validators(vals)
    recs = []
    for i in range(len(commit["signatures"])):
        v = ordered[i]
        recs.append(struct.pack("<32sQI4x", validator_pubkey(v), int(v["voting_power"]), len(validator_bytes(v))))
    while len(recs) < n_max:
        recs.append(struct.pack("<32sQI4x", DUMMY_PUBLIC_KEY, 0, 46))
    return recs


def _hash_proof(leaves, proofs, idx):
    assert len(leaves[idx]) == 34 and len(proofs[idx]) == 4
    return struct.pack("<34s2x", leaves[idx]) + b"".join(proofs[idx])


def _head(kind, n_max, nb_val, nb_trusted, rnd, header, leaves, proofs, aux_proof, last_block_id_proof):
    enc_chain = leaves[1]
    chain = b"".join(proofs[1]) + struct.pack("<I52s", len(enc_chain), enc_chain.ljust(52, b"\x00"))
    hv = leaves[2]
    # decode varint height back (leaf = 08 <varint>)
    hval, shift = 0, 0
    for b in hv[1:]:
        hval |= (b & 0x7F) << shift
        shift += 7
    height_p = b"".join(proofs[2]) + struct.pack("<I4xQ", len(hv), hval)
    out = struct.pack("<IIIIIIQ32s", MAGIC, kind, n_max, nb_val, nb_trusted, 0, rnd, header)
    out += chain + height_p + _hash_proof(leaves, proofs, 7) + aux_proof + last_block_id_proof
    assert len(out) == 920, len(out)
    return out


def skip_inputs(src, n_max, trusted_block, trusted_hash, target_block):
    """REF circuits/input/mod.rs:425-523.  Returns the blob (bytes)."""
    tv, gv = src.validators(trusted_block), src.validators(target_block)
    assert len(tv) <= n_max and len(gv) <= n_max, "validator set larger than VALIDATOR_SET_SIZE_MAX"
    tsh, gsh = src.signed_header(trusted_block), src.signed_header(target_block)
    assert header_hash(tsh["header"]) == trusted_hash, "Trusted header hash doesn't pass sanity check"
    gl_, tl_ = header_leaves(gsh["header"]), header_leaves(tsh["header"])
    groot, gproofs = merkle_proofs(gl_)
    troot, tproofs = merkle_proofs(tl_)
    head = _head(KIND_SKIP, n_max, len(gv), len(tv), int(gsh["commit"]["round"]), groot, gl_, gproofs,
                 _hash_proof(tl_, tproofs, 7), bytes(200))
    recs = validator_data_from_block(gv, gsh, n_max)
    fields = hash_fields_from_block(tv, tsh["commit"], n_max)
    return head + b"".join(recs) + b"".join(fields)


def step_inputs(src, n_max, prev_block, prev_hash):
    """REF circuits/input/mod.rs:316-423."""
    psh, nsh = src.signed_header(prev_block), src.signed_header(prev_block + 1)
    assert header_hash(psh["header"]) == prev_hash, "Prev header hash doesn't pass sanity check"
    nv = src.validators(prev_block + 1)
    assert len(nv) <= n_max
    nl, pl = header_leaves(nsh["header"]), header_leaves(psh["header"])
    nroot, nproofs = merkle_proofs(nl)
    proot, pproofs = merkle_proofs(pl)
    assert nl[4][2:34] == bytes.fromhex(nsh["header"]["last_block_id"]["hash"])
    lbi = struct.pack("<72s", nl[4]) + b"".join(nproofs[4])
    head = _head(KIND_STEP, n_max, len(nv), 0, int(nsh["commit"]["round"]), nroot, nl, nproofs,
                 _hash_proof(pl, pproofs, 8), lbi)
    return head + b"".join(validator_data_from_block(nv, nsh, n_max))


def skip_public_input(trusted_block, trusted_hash, target_block):
    return struct.pack(">Q", trusted_block) + trusted_hash + struct.pack(">Q", target_block)


def step_public_input(prev_block, prev_hash):
    return struct.pack(">Q", prev_block) + prev_hash


# ---- synthetic chain ----
def synthetic_source(seed=0, n_validators=128, chain_id="celestia", trusted_height=999_000, target_height=1_000_000,
                     absent_frac=0.0, rnd=0, step=False):
    """Deterministic chain with `n_validators` signers: returns (MemorySource, trusted_height, target_height).
    With step=True target_height = trusted_height + 1 and the headers are linked (last_block_id,
    next_validators_hash)."""
    from nacl.signing import SigningKey

    S = struct.pack("<Q", seed)
    keys = [SigningKey(sha256(b"tmx/val" + S + struct.pack("<I", i))) for i in range(n_validators)]
    vals = []
    for i, k in enumerate(keys):
        pk = bytes(k.verify_key)
        vals.append({"address": sha256(pk)[:20].hex().upper(),
                     "pub_key": {"type": "tendermint/PubKeyEd25519", "value": base64.b64encode(pk).decode()},
                     "voting_power": str(10 ** (i % 7) + i), "proposer_priority": "0", "_key": i})
    vals = sort_validators(vals)
    vh = validators_hash(vals)
    if step:
        target_height = trusted_height + 1
    hx = lambda tag, f: sha256(b"tmx/hdr" + S + tag + struct.pack("<I", f)).hex().upper()

    def mk_header(height, tag, last_block_id):
        return {"version": {"block": "11", "app": "1"}, "chain_id": chain_id, "height": str(height),
                "time": "2024-01-01T00:00:00.123456789Z", "last_block_id": last_block_id,
                "last_commit_hash": hx(tag, 5), "data_hash": hx(tag, 6), "validators_hash": vh.hex().upper(),
                "next_validators_hash": vh.hex().upper(), "consensus_hash": hx(tag, 9), "app_hash": hx(tag, 10),
                "last_results_hash": hx(tag, 11), "evidence_hash": hx(tag, 12),
                "proposer_address": vals[0]["address"]}

    def mk_commit(header, height):
        bid = {"hash": header_hash(header).hex().upper(), "parts": {"total": 1, "hash": sha256(b"tmx/psh" + S).hex().upper()}}
        sigs = []
        for j, v in enumerate(vals):
            absent = absent_frac > 0 and (sha256(b"tmx/abs" + S + struct.pack("<I", j))[0] / 256.0) < absent_frac
            ts = f"2024-01-01T00:00:01.{j + 1:09d}Z"
            if absent:
                sigs.append({"block_id_flag": 1, "validator_address": "", "timestamp": "0001-01-01T00:00:00Z", "signature": None})
                continue
            msg = sign_bytes(chain_id, height, rnd, bid, ts)
            sig = keys[v["_key"]].sign(msg).signature
            sigs.append({"block_id_flag": 2, "validator_address": v["address"], "timestamp": ts,
                         "signature": base64.b64encode(sig).decode()})
        return {"height": str(height), "round": rnd, "block_id": bid, "signatures": sigs}

    genesis_bid = {"hash": hx(b"t", 4), "parts": {"total": 1, "hash": hx(b"t", 40)}}
    th = mk_header(trusted_height, b"t", genesis_bid)
    tc = mk_commit(th, trusted_height)
    last = tc["block_id"] if step else {"hash": hx(b"g", 4), "parts": {"total": 1, "hash": hx(b"g", 40)}}
    gh = mk_header(target_height, b"g", last)
    gc = mk_commit(gh, target_height)
    pub = [{k: v for k, v in x.items() if k != "_key"} for x in vals]
    src = MemorySource({trusted_height: {"header": th, "commit": tc}, target_height: {"header": gh, "commit": gc}},
                       {trusted_height: pub, target_height: pub})
    return src, trusted_height, target_height


class SignedBlockSource:
    """<root>/<height>/signed_block.json (header + commit + validator_set in one file); the reference has no
    reader for it (REF circuits/input/tendermint_utils.rs:82-114 types are unused) but the data is genuine."""

    def __init__(self, root):
        self.root = root

    def _load(self, height):
        with open(os.path.join(self.root, str(height), "signed_block.json")) as f:
            return json.load(f)["result"]

    def signed_header(self, height):
        r = self._load(height)
        return {"header": r["header"], "commit": r["commit"]}

    def validators(self, height):
        return self._load(height)["validator_set"]["validators"]


# ---- operator-side search ----
def is_valid_skip(src, start_block, target_block):
    """REF circuits/input/tendermint_utils.rs:444-482.  `validator_address()` is Some for block_id_flag 2 (commit) and
    3 (nil), None for 1 (absent); the comparison is done in f64 exactly like the reference."""
    start, target = src.validators(start_block), src.validators(target_block)
    sigs = src.signed_header(target_block)["commit"]["signatures"]
    threshold = 1.0 / 3.0
    total = sum(int(v["voting_power"]) for v in target)
    by_addr = {}
    for v in target:
        by_addr.setdefault(v["address"].upper(), v)
    shared, idx = 0, 0
    while float(total) * threshold > float(shared) and idx < len(start):
        tv = by_addr.get(start[idx]["address"].upper())
        if tv is not None:
            for s in sigs:
                if int(s["block_id_flag"]) in (2, 3) and s["validator_address"].upper() == tv["address"].upper():
                    shared += int(tv["voting_power"])
        idx += 1
    return float(total) * threshold <= float(shared)


def find_block_to_request(src, start_block, max_end_block):
    """REF circuits/input/mod.rs:158-186."""
    cur = max_end_block
    while True:
        if cur - start_block == 1:
            return cur
        if is_valid_skip(src, start_block, cur):
            return cur
        cur = (cur + start_block) // 2
