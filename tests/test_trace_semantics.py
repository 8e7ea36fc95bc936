"""The witness tables read against the SPECIFICATIONS by tests/spec_checker.py (plain Python: FIPS 180-4 round functions, RFC 8032
affine arithmetic, hashlib), independently of the constraint code: every SHA-256 / SHA-512 chunk of the schedule shows the right
round states, schedule words, chaining values and digests for the right MESSAGE (validator leaves, tree nodes with the
promote-left rule, header proofs, R || A || M), and every Ed25519 slot the running value of [s]B + [h](-A), in-range exact
multiplication gadgets and a result equal to R.  Here on the CPU oracle's tables; tests/test_gpu_witness.py runs the same checker
on the tables the GPU kernels produce."""
import json
import os
import re
import struct

import pytest

import spec_checker as sc

HERE = os.path.dirname(os.path.abspath(__file__))
DUMMY_PK = bytes.fromhex("3b6a27bcceb6a42d62a3a8d02a6f0d73653215771de243a63ac048a18b59da29")
with open(os.path.join(os.path.dirname(HERE), "tendermintx_b200", "csrc", "dummy_sig.inc")) as _f:
    DUMMY_SIG = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-f]{2})", _f.read()))
assert len(DUMMY_SIG) == 64


def varint9(v):
    groups = [(v >> (7 * k)) & 0x7F for k in range(9)]
    last = max([k for k in range(9) if groups[k]] or [0])
    return bytes(g | (0x80 if k < last else 0) for k, g in enumerate(groups))


def schedule(blob):
    """(SHA-256 messages in chunk order, SHA-512 messages per slot, Ed25519 triples per slot) of a blob, from the reference's
    definitions [REF circuits/builder/validator.rs:185-252, shared.rs:43-65,169-207, verify.rs:180-222]."""
    import hashlib

    magic, kind, n_max, nb_val, nb_tr = struct.unpack_from("<IIIII", blob, 0)
    head = 920
    vals = [blob[head + 240 * i: head + 240 * (i + 1)] for i in range(n_max)]
    tfs = [blob[head + 240 * n_max + 48 * i: head + 240 * n_max + 48 * (i + 1)] for i in range(n_max)] if kind == 1 else []
    np2 = 1
    while np2 < n_max:
        np2 *= 2
    msgs = []

    def validator_set(entries, nb):
        nodes, en = [], []
        for i, (pk, power, blen) in enumerate(entries):
            m = b"\x00" + (b"\x0a\x22\x0a\x20" + pk + b"\x10" + varint9(power))[:blen]
            msgs.append(m)
            nodes.append(hashlib.sha256(m).digest())
            en.append(i < nb)
        nodes += [bytes(32)] * (np2 - n_max)
        en += [False] * (np2 - n_max)
        while len(nodes) > 1:
            nn, ne = [], []
            for i in range(0, len(nodes), 2):
                m = b"\x01" + nodes[i] + nodes[i + 1]
                msgs.append(m)
                nn.append(hashlib.sha256(m).digest() if en[i] and en[i + 1] else nodes[i])
                ne.append(en[i])
            nodes, en = nn, ne
        return nodes[0]

    roots = []
    if kind == 1:
        roots.append(validator_set([(t[:32], struct.unpack_from("<Q", t, 32)[0], struct.unpack_from("<I", t, 40)[0]) for t in tfs], nb_tr))
    roots.append(validator_set([(v[:32], struct.unpack_from("<Q", v, 224)[0], struct.unpack_from("<I", v, 232)[0]) for v in vals], nb_val))

    def proof(leaf, aunts, index):
        msgs.append(leaf)
        cur = hashlib.sha256(leaf).digest()
        for j in range(4):
            a = aunts[32 * j: 32 * j + 32]
            m = b"\x01" + (a + cur if (index >> j) & 1 else cur + a)
            msgs.append(m)
            cur = hashlib.sha256(m).digest()
        return cur

    chain_aunts, chain_len = blob[64:192], struct.unpack_from("<I", blob, 192)[0]
    chain_leaf = b"\x00" + blob[196:196 + chain_len]
    height_aunts, height_len, height = blob[248:376], struct.unpack_from("<I", blob, 376)[0], struct.unpack_from("<Q", blob, 384)[0]
    height_leaf = b"\x00" + (b"\x08" + varint9(height))[:height_len]
    vh_leaf, vh_aunts = b"\x00" + blob[392:426], blob[428:556]
    aux_leaf, aux_aunts = b"\x00" + blob[556:590], blob[592:720]
    lb_leaf, lb_aunts = b"\x00" + blob[720:792], blob[792:920]
    header = blob[32:64]
    reached = []
    if kind == 1:
        reached.append(proof(aux_leaf, aux_aunts, 7))
        reached += [proof(vh_leaf, vh_aunts, 7), proof(chain_leaf, chain_aunts, 1), proof(height_leaf, height_aunts, 2)]
        assert reached[1:] == [header] * 3 and aux_leaf[3:] == roots[0] and vh_leaf[3:] == roots[1]
    else:
        reached += [proof(vh_leaf, vh_aunts, 7), proof(chain_leaf, chain_aunts, 1), proof(height_leaf, height_aunts, 2),
                    proof(lb_leaf, lb_aunts, 4), proof(aux_leaf, aux_aunts, 8)]
        assert reached[:4] == [header] * 4 and vh_leaf[3:] == roots[0] == aux_leaf[3:]
    triples = []
    for v in vals:
        if v[236]:
            ln = struct.unpack_from("<I", v, 220)[0]
            triples.append((v[:32], v[32:96], v[96:96 + ln]))
        else:
            triples.append((DUMMY_PK, DUMMY_SIG, bytes(32)))
    return msgs, [t[1][:32] + t[0] + t[2] for t in triples], triples


def check_tables(tabs, blob):
    msgs, msgs512, triples = schedule(blob)
    sc.check_sha256_table(tabs[0], msgs)
    sc.check_sha512_table(tabs[1], msgs512)
    sc.check_ed25519_table(tabs[2], triples)
    sc.check_mul_gadgets_exact(tabs[2], range(0, 256 * len(triples), 37))


@pytest.mark.parametrize("name", ["skip_3000_3100_n4", "step_10500_n4_with_dummy"])
def test_oracle_tables_encode_the_specified_computations(oracle, name):
    with open(os.path.join(HERE, "golden", "fixture_vectors.json")) as f:
        c = {x["name"]: x for x in json.load(f)["cases"]}[name]
    blob = bytes.fromhex(c["blob"])
    check_tables(oracle.build_traces(blob), blob)
