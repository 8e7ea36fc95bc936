// Host-side input assembly: the fixture mode of the reference's InputDataFetcher, in C++.
//
// Mirrors circuits/input/mod.rs (get_signed_header_from_number :188-217, get_validator_set_from_number :219-241,
// fetch_validator_result :243-282, get_skip_inputs :425-523, get_step_inputs :316-423), circuits/input/conversion.rs
// (get_validator_data_from_block :59-137, validator_hash_field_from_block :139-178), and
// circuits/input/tendermint_utils.rs (header leaf encoding :374-393, RFC-6962 Merkle proofs :294-349) together
// with the tendermint-rs encodings they call (SignedVote::sign_bytes, Info::hash_bytes, validator::Set ordering).
// Output: the packed off-chain blob of include/tmx_types.h, i.e. what the async hints hand to the circuit
// [REF circuits/skip.rs:64-101, circuits/step.rs:56-88].  RPC mode is out of scope (no network); the JSON shapes
// are the RPC ones, so a caller that has fetched responses itself can drop them in a directory.
// Not inherited on purpose: the TENDERMINT_RPC_URL requirement and the hard-wired relative fixture path
// [REF mod.rs:81,88]; the host-side signature sanity check [REF conversion.rs:48] (the Ed25519 kernel re-verifies
// every signature and the proof is refused with TMX_E_UNSAT if one fails).
#include "ctx.cuh"
#include "witness.cuh"
#include <cstdio>
#include <cstring>
#include <ctime>
#include <memory>
#include <array>
#include <stdexcept>
#include <algorithm>

namespace tmx {
namespace {

// ---------------------------------------------------------------- minimal JSON (RPC responses are plain UTF-8)
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    std::string s;  // string value, or the raw text of a number
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json* get(const char* key) const {
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct JsonParser {
    const char* p;
    const char* end;
    bool ok = true;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool lit(const char* w) {
        size_t n = strlen(w);
        if ((size_t)(end - p) >= n && !memcmp(p, w, n)) { p += n; return true; }
        return false;
    }
    std::string str() {
        std::string out;
        p++;  // opening quote
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': p += 4; out += '?'; break;  // not needed for RPC payloads
                    default: out += *p;
                }
                p++;
            } else
                out += *p++;
        }
        if (p >= end) ok = false; else p++;
        return out;
    }
    Json value() {
        Json j;
        ws();
        if (p >= end) { ok = false; return j; }
        if (*p == '{') {
            j.kind = Json::Obj;
            p++;
            ws();
            if (p < end && *p == '}') { p++; return j; }
            while (ok) {
                ws();
                if (p >= end || *p != '"') { ok = false; break; }
                std::string k = str();
                ws();
                if (p >= end || *p != ':') { ok = false; break; }
                p++;
                j.obj.emplace_back(k, value());
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                ok = false;
            }
        } else if (*p == '[') {
            j.kind = Json::Arr;
            p++;
            ws();
            if (p < end && *p == ']') { p++; return j; }
            while (ok) {
                j.arr.push_back(value());
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                ok = false;
            }
        } else if (*p == '"') {
            j.kind = Json::Str;
            j.s = str();
        } else if (lit("true")) {
            j.kind = Json::Bool; j.b = true;
        } else if (lit("false")) {
            j.kind = Json::Bool;
        } else if (lit("null")) {
            j.kind = Json::Null;
        } else {
            j.kind = Json::Num;
            const char* s = p;
            while (p < end && (strchr("+-.eE", *p) || (*p >= '0' && *p <= '9'))) p++;
            if (p == s) ok = false;
            j.s.assign(s, p);
        }
        return j;
    }
};

struct InputError {
    int code;
    std::string msg;
};
[[noreturn]] void bad(const std::string& m) { throw InputError{TMX_E_INPUT, m}; }

const Json& need(const Json& j, const char* key) {
    const Json* v = j.get(key);
    if (!v) bad(std::string("missing JSON field '") + key + "'");
    return *v;
}
uint64_t as_u64(const Json& j) {  // RPC encodes 64-bit integers as strings, small ones as numbers
    if (j.kind != Json::Str && j.kind != Json::Num) bad("expected an integer");
    return strtoull(j.s.c_str(), nullptr, 10);
}
std::string as_str(const Json& j) {
    if (j.kind == Json::Null) return "";
    if (j.kind != Json::Str) bad("expected a string");
    return j.s;
}

Json load_json(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw InputError{TMX_E_IO, "cannot open " + path};
    std::string buf;
    char tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
    fclose(f);
    JsonParser ps{buf.data(), buf.data() + buf.size()};
    Json j = ps.value();
    if (!ps.ok) bad("malformed JSON in " + path);
    return j;
}

// ---------------------------------------------------------------- byte helpers
typedef std::vector<uint8_t> Bytes;
Bytes from_hex(const std::string& h) {
    if (h.size() % 2) bad("odd-length hex string");
    Bytes out(h.size() / 2);
    auto nib = [](char c) -> int {
        if (c >= '0' && c <= '9') return c - '0';
        if (c >= 'a' && c <= 'f') return c - 'a' + 10;
        if (c >= 'A' && c <= 'F') return c - 'A' + 10;
        bad("bad hex digit");
    };
    for (size_t i = 0; i < out.size(); i++) out[i] = (uint8_t)(nib(h[2 * i]) << 4 | nib(h[2 * i + 1]));
    return out;
}
Bytes from_base64(const std::string& s) {
    Bytes out;
    uint32_t acc = 0;
    int bits = 0;
    for (char c : s) {
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+') v = 62;
        else if (c == '/') v = 63;
        else if (c == '=') break;
        else bad("bad base64 character");
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)(acc >> bits));
        }
    }
    return out;
}
void put_varint(Bytes& b, uint64_t v) {
    do {
        uint8_t x = v & 0x7F;
        v >>= 7;
        b.push_back(x | (v ? 0x80 : 0));
    } while (v);
}
void put(Bytes& b, const Bytes& x) { b.insert(b.end(), x.begin(), x.end()); }
Bytes len_prefixed(uint8_t tag, const Bytes& x) {
    Bytes b{tag};
    put_varint(b, x.size());
    put(b, x);
    return b;
}
Bytes field1_bytes(const Bytes& x) { return x.empty() ? Bytes{} : len_prefixed(0x0a, x); }

// RFC 3339 "2023-09-07T14:22:28.360824457Z" -> (seconds, nanos)
void parse_time(const std::string& s, int64_t* secs, int64_t* nanos) {
    int Y, M, D, h, m, sec;
    if (sscanf(s.c_str(), "%d-%d-%dT%d:%d:%d", &Y, &M, &D, &h, &m, &sec) != 6) bad("bad timestamp " + s);
    // days from civil (proleptic Gregorian)
    int y = Y - (M <= 2);
    int era = (y >= 0 ? y : y - 399) / 400;
    unsigned yoe = (unsigned)(y - era * 400);
    unsigned doy = (153 * (M + (M > 2 ? -3 : 9)) + 2) / 5 + D - 1;
    unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    int64_t days = (int64_t)era * 146097 + (int64_t)doe - 719468;
    *secs = days * 86400 + h * 3600 + m * 60 + sec;
    *nanos = 0;
    size_t dot = s.find('.');
    if (dot != std::string::npos) {
        std::string frac;
        for (size_t i = dot + 1; i < s.size() && s[i] >= '0' && s[i] <= '9'; i++) frac += s[i];
        frac.resize(9, '0');
        *nanos = strtoll(frac.c_str(), nullptr, 10);
    }
}
Bytes enc_timestamp(const std::string& s) {
    int64_t secs, nanos;
    parse_time(s, &secs, &nanos);
    Bytes b;
    if (secs) { b.push_back(0x08); put_varint(b, (uint64_t)secs); }
    if (nanos) { b.push_back(0x10); put_varint(b, (uint64_t)nanos); }
    return b;
}

// ---------------------------------------------------------------- SHA-256 / Merkle on the host
void sha256_host(const uint8_t* msg, size_t len, uint8_t out[32]) {
    std::vector<uint8_t> buf(((len + 9 + 63) / 64) * 64);
    const int nb = sha256_pad_blocks(msg, (int)len, buf.data());
    uint32_t st[8];
    for (int i = 0; i < 8; i++) st[i] = iv256(i);
    Sha256Hist hs;
    for (int b = 0; b < nb; b++) sha256_compress_hist(st, buf.data() + 64 * b, &hs, st);
    sha256_state_to_bytes(st, out);
}
typedef std::array<uint8_t, 32> Hash;
Hash leaf_hash(const Bytes& x) {
    Bytes m{0};
    put(m, x);
    Hash h;
    sha256_host(m.data(), m.size(), h.data());
    return h;
}
Hash inner_hash(const Hash& l, const Hash& r) {
    uint8_t m[65];
    m[0] = 1;
    memcpy(m + 1, l.data(), 32);
    memcpy(m + 33, r.data(), 32);
    Hash h;
    sha256_host(m, 65, h.data());
    return h;
}
size_t split_point(size_t n) {  // REF tendermint_utils.rs:338-349
    size_t k = 1;
    while (k * 2 < n) k *= 2;
    return k;
}
// root and per-leaf aunts (leaf upwards), REF tendermint_utils.rs:294-336
Hash merkle(const std::vector<Bytes>& items, size_t lo, size_t hi, std::vector<std::vector<Hash>>& aunts) {
    if (hi - lo == 1) return leaf_hash(items[lo]);
    const size_t k = split_point(hi - lo);
    Hash l = merkle(items, lo, lo + k, aunts), r = merkle(items, lo + k, hi, aunts);
    for (size_t i = lo; i < lo + k; i++) aunts[i].push_back(r);
    for (size_t i = lo + k; i < hi; i++) aunts[i].push_back(l);
    return inner_hash(l, r);
}

// ---------------------------------------------------------------- Tendermint encodings
Bytes enc_block_id(const Json& bid) {
    const Bytes h = from_hex(as_str(need(bid, "hash")));
    const Json& parts = need(bid, "parts");
    Bytes psh;
    const uint64_t total = as_u64(need(parts, "total"));
    if (total) { psh.push_back(0x08); put_varint(psh, total); }
    const Bytes ph = from_hex(as_str(need(parts, "hash")));
    if (!ph.empty()) put(psh, len_prefixed(0x12, ph));
    Bytes out;
    if (!h.empty()) put(out, len_prefixed(0x0a, h));
    put(out, len_prefixed(0x12, psh));
    return out;
}
// the 14 header leaves in the order of REF tendermint_utils.rs:374-393
std::vector<Bytes> header_leaves(const Json& h) {
    std::vector<Bytes> L;
    const Json& ver = need(h, "version");
    Bytes v{0x08};
    put_varint(v, as_u64(need(ver, "block")));
    if (ver.get("app") && as_u64(*ver.get("app"))) { v.push_back(0x10); put_varint(v, as_u64(*ver.get("app"))); }
    L.push_back(v);
    const std::string chain = as_str(need(h, "chain_id"));
    L.push_back(field1_bytes(Bytes(chain.begin(), chain.end())));
    Bytes hv{0x08};
    put_varint(hv, as_u64(need(h, "height")));
    L.push_back(hv);
    L.push_back(enc_timestamp(as_str(need(h, "time"))));
    L.push_back(enc_block_id(need(h, "last_block_id")));
    for (const char* k : {"last_commit_hash", "data_hash", "validators_hash", "next_validators_hash", "consensus_hash", "app_hash",
                          "last_results_hash", "evidence_hash", "proposer_address"})
        L.push_back(field1_bytes(from_hex(as_str(need(h, k)))));
    return L;
}
Bytes validator_bytes(const Bytes& pk, uint64_t power) {  // Info::hash_bytes
    Bytes b{0x0a, 0x22, 0x0a, 0x20};
    put(b, pk);
    if (power) { b.push_back(0x10); put_varint(b, power); }
    return b;
}
// CanonicalVote sign bytes (SignedVote::sign_bytes as used at REF conversion.rs:34-39)
Bytes sign_bytes(const std::string& chain_id, uint64_t height, uint64_t round, const Json& block_id, const std::string& ts) {
    Bytes body{0x08, 0x02, 0x11};
    for (int i = 0; i < 8; i++) body.push_back((uint8_t)(height >> (8 * i)));
    if (round) {
        body.push_back(0x19);
        for (int i = 0; i < 8; i++) body.push_back((uint8_t)(round >> (8 * i)));
    }
    const Bytes h = from_hex(as_str(need(block_id, "hash")));
    const Json& parts = need(block_id, "parts");
    Bytes psh{0x08};
    put_varint(psh, as_u64(need(parts, "total")));
    put(psh, len_prefixed(0x12, from_hex(as_str(need(parts, "hash")))));
    Bytes cb = len_prefixed(0x0a, h);
    put(cb, len_prefixed(0x12, psh));
    put(body, len_prefixed(0x22, cb));
    put(body, len_prefixed(0x2a, enc_timestamp(ts)));
    put(body, len_prefixed(0x32, Bytes(chain_id.begin(), chain_id.end())));
    Bytes out;
    put_varint(out, body.size());
    put(out, body);
    return out;
}

struct Validator {
    Bytes address, pubkey;
    uint64_t power;
};

// ---------------------------------------------------------------- fixture source (REF mod.rs:188-282)
struct FixtureSource {
    std::string root;
    Json signed_header(uint64_t h) const {
        Json j = load_json(root + "/" + std::to_string(h) + "/commit.json");
        return need(need(j, "result"), "signed_header");
    }
    std::vector<Validator> validators(uint64_t h) const {
        std::vector<Validator> out;
        for (int page = 1;; page++) {
            Json j = load_json(root + "/" + std::to_string(h) + "/validators_" + std::to_string(page) + ".json");
            const Json& r = need(j, "result");
            for (const Json& v : need(r, "validators").arr) {
                Validator x;
                x.address = from_hex(as_str(need(v, "address")));
                x.pubkey = from_base64(as_str(need(need(v, "pub_key"), "value")));
                if (x.pubkey.size() != 32) bad("validator public key is not 32 bytes");
                x.power = as_u64(need(v, "voting_power"));
                out.push_back(x);
            }
            if (out.size() >= as_u64(need(r, "total")) || need(r, "validators").arr.empty()) return out;
        }
    }
};

struct HeaderProofs {
    std::vector<Bytes> leaves;
    std::vector<std::vector<Hash>> aunts;
    Hash root;
};
HeaderProofs prove_header(const Json& header) {
    HeaderProofs hp;
    hp.leaves = header_leaves(header);
    hp.aunts.resize(hp.leaves.size());
    hp.root = merkle(hp.leaves, 0, hp.leaves.size(), hp.aunts);
    return hp;
}
void fill_hash_proof(tmx_hash_proof* out, const HeaderProofs& hp, size_t idx) {
    if (hp.leaves[idx].size() != 34 || hp.aunts[idx].size() != 4) bad("header leaf has an unexpected shape");
    memcpy(out->leaf, hp.leaves[idx].data(), 34);
    for (int i = 0; i < 4; i++) memcpy(out->aunts[i], hp.aunts[idx][i].data(), 32);
}

const uint8_t DUMMY_PK[32] = {0x3b, 0x6a, 0x27, 0xbc, 0xce, 0xb6, 0xa4, 0x2d, 0x62, 0xa3, 0xa8, 0xd0, 0x2a, 0x6f, 0x0d, 0x73,
                              0x65, 0x32, 0x15, 0x77, 0x1d, 0xe2, 0x43, 0xa6, 0x3a, 0xc0, 0x48, 0xa1, 0x8b, 0x59, 0xda, 0x29};
#include "dummy_sig.inc"

// REF conversion.rs:59-137
void validator_data_from_block(const std::vector<Validator>& vals, const Json& sh, uint32_t n_max, tmx_validator* out) {
    const Json& header = need(sh, "header");
    const Json& commit = need(sh, "commit");
    const std::string chain = as_str(need(header, "chain_id"));
    const uint64_t height = as_u64(need(commit, "height")), round = as_u64(need(commit, "round"));
    const auto& sigs = need(commit, "signatures").arr;
    size_t k = 0;
    for (size_t i = 0; i < sigs.size(); i++) {
        if (i >= vals.size()) bad("more commit signatures than validators");
        if (k >= n_max) bad("validator set larger than VALIDATOR_SET_SIZE_MAX");
        const Validator& v = vals[i];
        tmx_validator& r = out[k++];
        memset(&r, 0, sizeof r);
        memcpy(r.pubkey, v.pubkey.data(), 32);
        r.voting_power = v.power;
        r.validator_byte_length = (uint32_t)validator_bytes(v.pubkey, v.power).size();
        if (as_u64(need(sigs[i], "block_id_flag")) == 2) {
            const Bytes msg = sign_bytes(chain, height, round, need(commit, "block_id"), as_str(need(sigs[i], "timestamp")));
            const Bytes sig = from_base64(as_str(need(sigs[i], "signature")));
            if (msg.size() > TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX || sig.size() != 64) bad("sign-bytes / signature of unexpected size");
            memcpy(r.message, msg.data(), msg.size());
            r.message_byte_length = (uint32_t)msg.size();
            memcpy(r.sig_r, sig.data(), 32);
            memcpy(r.sig_s, sig.data() + 32, 32);
            r.is_signed = 1;
        } else {
            memcpy(r.sig_r, DUMMY_SIGNATURE, 32);
            memcpy(r.sig_s, DUMMY_SIGNATURE + 32, 32);
            r.message_byte_length = 32;
        }
    }
    for (; k < n_max; k++) {
        tmx_validator& r = out[k];
        memset(&r, 0, sizeof r);
        memcpy(r.pubkey, DUMMY_PK, 32);
        memcpy(r.sig_r, DUMMY_SIGNATURE, 32);
        memcpy(r.sig_s, DUMMY_SIGNATURE + 32, 32);
        r.message_byte_length = 32;
        r.validator_byte_length = TMX_VALIDATOR_BYTE_LENGTH_MAX;
    }
}

// REF conversion.rs:139-178 (validator::Set ordering: power descending, address ascending)
void hash_fields_from_block(std::vector<Validator> vals, const Json& commit, uint32_t n_max, tmx_hash_field* out) {
    std::stable_sort(vals.begin(), vals.end(), [](const Validator& a, const Validator& b) {
        if (a.power != b.power) return a.power > b.power;
        return a.address < b.address;
    });
    const size_t nsig = need(commit, "signatures").arr.size();
    size_t k = 0;
    for (size_t i = 0; i < nsig; i++) {
        if (i >= vals.size()) bad("more commit signatures than validators");
        if (k >= n_max) bad("validator set larger than VALIDATOR_SET_SIZE_MAX");
        tmx_hash_field& f = out[k++];
        memset(&f, 0, sizeof f);
        memcpy(f.pubkey, vals[i].pubkey.data(), 32);
        f.voting_power = vals[i].power;
        f.validator_byte_length = (uint32_t)validator_bytes(vals[i].pubkey, vals[i].power).size();
    }
    for (; k < n_max; k++) {
        tmx_hash_field& f = out[k];
        memset(&f, 0, sizeof f);
        memcpy(f.pubkey, DUMMY_PK, 32);
        f.validator_byte_length = TMX_VALIDATOR_BYTE_LENGTH_MAX;
    }
}

void fill_head_common(tmx_offchain_head* h, uint32_t kind, uint32_t n_max, const Json& sh, const HeaderProofs& hp, size_t nb_val) {
    memset(h, 0, sizeof *h);
    h->magic = TMX_BLOB_MAGIC;
    h->kind = kind;
    h->n_max = n_max;
    h->nb_validators = (uint32_t)nb_val;
    h->round = as_u64(need(need(sh, "commit"), "round"));
    memcpy(h->header, hp.root.data(), 32);
    const Bytes& chain = hp.leaves[TMX_CHAIN_ID_INDEX];
    if (chain.size() > TMX_PROTOBUF_CHAIN_ID_SIZE_BYTES) bad("chain id longer than 50 characters");
    for (int i = 0; i < 4; i++) memcpy(h->chain_id_proof.aunts[i], hp.aunts[TMX_CHAIN_ID_INDEX][i].data(), 32);
    h->chain_id_proof.enc_chain_id_byte_length = (uint32_t)chain.size();
    memcpy(h->chain_id_proof.chain_id, chain.data(), chain.size());
    const Bytes& hl = hp.leaves[TMX_BLOCK_HEIGHT_INDEX];
    for (int i = 0; i < 4; i++) memcpy(h->height_proof.aunts[i], hp.aunts[TMX_BLOCK_HEIGHT_INDEX][i].data(), 32);
    h->height_proof.enc_height_byte_length = (uint32_t)hl.size();
    h->height_proof.height = as_u64(need(need(sh, "header"), "height"));
    fill_hash_proof(&h->validators_hash_proof, hp, TMX_VALIDATORS_HASH_INDEX);
}

int build_skip(const FixtureSource& src, uint32_t n_max, uint64_t trusted, const uint8_t trusted_hash[32], uint64_t target, uint8_t* blob) {
    const std::vector<Validator> tv = src.validators(trusted), gv = src.validators(target);
    if (tv.size() > n_max || gv.size() > n_max) bad("The validator set size of the trusted or target block is larger than the VALIDATOR_SET_SIZE_MAX.");
    const Json tsh = src.signed_header(trusted), gsh = src.signed_header(target);
    const HeaderProofs tp = prove_header(need(tsh, "header")), gp = prove_header(need(gsh, "header"));
    if (memcmp(tp.root.data(), trusted_hash, 32)) bad("Trusted header hash doesn't pass sanity check!");
    tmx_offchain_head* h = (tmx_offchain_head*)blob;
    fill_head_common(h, TMX_KIND_SKIP, n_max, gsh, gp, gv.size());
    h->nb_trusted = (uint32_t)tv.size();
    fill_hash_proof(&h->aux_hash_proof, tp, TMX_VALIDATORS_HASH_INDEX);
    tmx_validator* vals = (tmx_validator*)(blob + sizeof(tmx_offchain_head));
    validator_data_from_block(gv, gsh, n_max, vals);
    hash_fields_from_block(tv, need(tsh, "commit"), n_max, (tmx_hash_field*)(vals + n_max));
    return TMX_OK;
}

int build_step(const FixtureSource& src, uint32_t n_max, uint64_t prev, const uint8_t prev_hash[32], uint8_t* blob) {
    const Json psh = src.signed_header(prev), nsh = src.signed_header(prev + 1);
    const HeaderProofs pp = prove_header(need(psh, "header")), np = prove_header(need(nsh, "header"));
    if (memcmp(pp.root.data(), prev_hash, 32)) bad("Prev header hash doesn't pass sanity check");
    const std::vector<Validator> nv = src.validators(prev + 1);
    if (nv.size() > n_max) bad("The validator set size of the next block is larger than the VALIDATOR_SET_SIZE_MAX.");
    tmx_offchain_head* h = (tmx_offchain_head*)blob;
    fill_head_common(h, TMX_KIND_STEP, n_max, nsh, np, nv.size());
    fill_hash_proof(&h->aux_hash_proof, pp, TMX_NEXT_VALIDATORS_HASH_INDEX);
    const Bytes& lbi = np.leaves[TMX_LAST_BLOCK_ID_INDEX];
    if (lbi.size() != TMX_PROTOBUF_BLOCK_ID_SIZE_BYTES) bad("last_block_id leaf is not 72 bytes");
    memcpy(h->last_block_id_proof.leaf, lbi.data(), 72);
    for (int i = 0; i < 4; i++) memcpy(h->last_block_id_proof.aunts[i], np.aunts[TMX_LAST_BLOCK_ID_INDEX][i].data(), 32);
    validator_data_from_block(nv, nsh, n_max, (tmx_validator*)(blob + sizeof(tmx_offchain_head)));
    return TMX_OK;
}

// REF tendermint_utils.rs:444-482 (is_valid_skip): voting power of the target set held by validators of the start set
// that appear in the target commit with a validator address (tendermint-rs CommitSig::validator_address(): flag 2
// "commit" and flag 3 "nil" carry one, flag 1 "absent" does not), compared in f64 against one third of the target
// set's total power exactly as the reference does (1_f64 / 3_f64, >, <=).
bool is_valid_skip(const FixtureSource& src, uint64_t start_block, uint64_t target_block) {
    const std::vector<Validator> start = src.validators(start_block), target = src.validators(target_block);
    const Json sh = src.signed_header(target_block);
    const auto& sigs = need(need(sh, "commit"), "signatures").arr;
    const double threshold = 1.0 / 3.0;
    uint64_t total = 0, shared = 0;
    for (const Validator& v : target) total += v.power;
    size_t idx = 0;
    while ((double)total * threshold > (double)shared && idx < start.size()) {
        const Validator* tv = nullptr;
        for (const Validator& v : target)
            if (v.address == start[idx].address) {
                tv = &v;
                break;
            }
        if (tv) {
            for (const Json& sig : sigs) {
                const uint64_t flag = as_u64(need(sig, "block_id_flag"));
                if (flag != 2 && flag != 3) continue;
                if (from_hex(as_str(need(sig, "validator_address"))) == tv->address) shared += tv->power;
            }
        }
        idx++;
    }
    return (double)total * threshold <= (double)shared;
}

// REF input/mod.rs:158-186 (find_block_to_request): halve the distance until a skip is possible; start + 1 means "step"
uint64_t find_block_to_request(const FixtureSource& src, uint64_t start_block, uint64_t max_end_block) {
    if (max_end_block <= start_block) bad("find_block_to_request: max_end_block must be above start_block");
    uint64_t cur = max_end_block;
    for (;;) {
        if (cur - start_block == 1) return cur;
        if (is_valid_skip(src, start_block, cur)) return cur;
        cur = (cur + start_block) / 2;
    }
}

template <class Fn>
int guarded(Fn&& fn) {
    try {
        return fn();
    } catch (const InputError& e) {
        return fail(e.code, e.msg);
    } catch (const std::exception& e) {
        return fail(TMX_E_INPUT, e.what());
    }
}

}  // namespace
}  // namespace tmx

using namespace tmx;

extern "C" int tmx_header_hash_from_fixture(const char* fixture_dir, uint64_t block, uint8_t out[32]) {
    if (!fixture_dir || !out) return fail(TMX_E_INPUT, "tmx_header_hash_from_fixture: NULL argument");
    return guarded([&] {
        FixtureSource src{fixture_dir};
        const Json sh = src.signed_header(block);
        const HeaderProofs hp = prove_header(need(sh, "header"));
        memcpy(out, hp.root.data(), 32);
        return (int)TMX_OK;
    });
}

extern "C" int tmx_is_valid_skip_from_fixture(const char* fixture_dir, uint64_t start_block, uint64_t target_block, int* valid) {
    if (!fixture_dir || !valid) return fail(TMX_E_INPUT, "tmx_is_valid_skip_from_fixture: NULL argument");
    return guarded([&] {
        *valid = is_valid_skip(FixtureSource{fixture_dir}, start_block, target_block) ? 1 : 0;
        return (int)TMX_OK;
    });
}

extern "C" int tmx_find_block_to_request(const char* fixture_dir, uint64_t start_block, uint64_t max_end_block, uint64_t* block) {
    if (!fixture_dir || !block) return fail(TMX_E_INPUT, "tmx_find_block_to_request: NULL argument");
    return guarded([&] {
        *block = find_block_to_request(FixtureSource{fixture_dir}, start_block, max_end_block);
        return (int)TMX_OK;
    });
}

extern "C" int tmx_skip_inputs_from_fixture(const char* fixture_dir, uint32_t n_max, uint64_t trusted_block,
                                            const uint8_t trusted_hash[32], uint64_t target_block, uint8_t* blob, size_t cap) {
    if (!fixture_dir || !trusted_hash || !blob || n_max == 0) return fail(TMX_E_INPUT, "tmx_skip_inputs_from_fixture: bad arguments");
    if (cap < TMX_BLOB_SIZE(TMX_KIND_SKIP, n_max)) return fail(TMX_E_INPUT, "tmx_skip_inputs_from_fixture: blob buffer too small");
    return guarded([&] { return build_skip(FixtureSource{fixture_dir}, n_max, trusted_block, trusted_hash, target_block, blob); });
}

extern "C" int tmx_step_inputs_from_fixture(const char* fixture_dir, uint32_t n_max, uint64_t prev_block, const uint8_t prev_hash[32],
                                            uint8_t* blob, size_t cap) {
    if (!fixture_dir || !prev_hash || !blob || n_max == 0) return fail(TMX_E_INPUT, "tmx_step_inputs_from_fixture: bad arguments");
    if (cap < TMX_BLOB_SIZE(TMX_KIND_STEP, n_max)) return fail(TMX_E_INPUT, "tmx_step_inputs_from_fixture: blob buffer too small");
    return guarded([&] { return build_step(FixtureSource{fixture_dir}, n_max, prev_block, prev_hash, blob); });
}
