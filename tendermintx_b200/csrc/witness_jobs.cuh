// Job-level logic of the witness kernels: which message goes into which SHA-256 chunk slot, how the
// validator-set tree selects its nodes, what the SHA-512 / Ed25519 slot of validator i contains.  Each job is a
// `prepare` step (one thread: builds messages, runs the sequential compressions / ladders into a history
// buffer) and a `rows` step (one thread per trace row, coalesced column-major stores).  Host+device for the
// same reason as witness.cuh.
//
// SHA-256 schedule (fixed by circuit shape; mirrors the call order of REF circuits/builder/verify.rs:361-437,
// 224-334 and validator.rs:231-252):
//   for each validator set s (skip: 0 = trusted, 1 = target; step: 0 = target), base(s) = s * (N + 2 (Np - 1)):
//       slots base + i                                   leaf i (1 chunk)
//       slots base + N + 2 * (inner_off(l) + i) (+1)     inner node i of level l >= 1 (2 chunks)
//   then header proofs (leaf, 4 inner nodes): skip: trusted valhash, target valhash, chain id, height;
//   step: valhash, chain id, height, last block id (2-chunk leaf), prev next-valhash; then zero-block padding.
#pragma once
#include "witness.cuh"

namespace tmx {

static const uint8_t DUMMY_PUBLIC_KEY[32] = {0x3b, 0x6a, 0x27, 0xbc, 0xce, 0xb6, 0xa4, 0x2d, 0x62, 0xa3, 0xa8,
                                             0xd0, 0x2a, 0x6f, 0x0d, 0x73, 0x65, 0x32, 0x15, 0x77, 0x1d, 0xe2,
                                             0x43, 0xa6, 0x3a, 0xc0, 0x48, 0xa1, 0x8b, 0x59, 0xda, 0x29};
#include "dummy_sig.inc"
#if defined(__CUDACC__)
static __constant__ uint8_t d_DUMMY_PUBLIC_KEY[32] = {0x3b, 0x6a, 0x27, 0xbc, 0xce, 0xb6, 0xa4, 0x2d, 0x62, 0xa3, 0xa8,
                                                      0xd0, 0x2a, 0x6f, 0x0d, 0x73, 0x65, 0x32, 0x15, 0x77, 0x1d, 0xe2,
                                                      0x43, 0xa6, 0x3a, 0xc0, 0x48, 0xa1, 0x8b, 0x59, 0xda, 0x29};
static __constant__ uint8_t d_DUMMY_SIGNATURE[64];
#endif
TMX_HD uint8_t dummy_pk(int i) {
#if defined(__CUDA_ARCH__)
    return d_DUMMY_PUBLIC_KEY[i];
#else
    return DUMMY_PUBLIC_KEY[i];
#endif
}
TMX_HD uint8_t dummy_sig(int i) {
#if defined(__CUDA_ARCH__)
    return d_DUMMY_SIGNATURE[i];
#else
    return DUMMY_SIGNATURE[i];
#endif
}

struct WitnessArgs {
    const uint8_t* blob;
    uint32_t kind, n_max, np, log_np;
    gl* t256;
    size_t n256;
    gl* t512;
    size_t n512;
    gl* ted;
    size_t ned;
    uint8_t* nodes;    // [n_sets][2 * np][32] selected node values, level l at offset 2np - (2np >> l)
    uint8_t* node_en;  // [n_sets][2 * np]
    uint8_t* aux;      // see AUX_* offsets
};
// aux layout (bytes)
constexpr size_t AUX_SET_ROOT = 0;      // [2][32] computed validators hash per set
constexpr size_t AUX_PROOF_ROOT = 64;   // [5][32] root reached by each header proof
constexpr size_t AUX_SIG_OK = 224;      // [n_max] 1 = signature equation holds
TMX_HD size_t aux_bytes(uint32_t n_max) { return AUX_SIG_OK + n_max; }

TMX_HD const tmx_offchain_head* blob_head(const uint8_t* blob) { return (const tmx_offchain_head*)blob; }
TMX_HD const tmx_validator* blob_validators(const uint8_t* blob) { return (const tmx_validator*)(blob + sizeof(tmx_offchain_head)); }
TMX_HD const tmx_hash_field* blob_hash_fields(const uint8_t* blob, uint32_t n_max) {
    return (const tmx_hash_field*)(blob + sizeof(tmx_offchain_head) + (size_t)n_max * sizeof(tmx_validator));
}
TMX_HD uint32_t n_sets(uint32_t kind) { return kind == TMX_KIND_SKIP ? 2 : 1; }
TMX_HD size_t set_chunks(uint32_t n_max, uint32_t np) { return (size_t)n_max + 2 * ((size_t)np - 1); }
TMX_HD size_t level_off(uint32_t np, uint32_t l) { return 2 * (size_t)np - ((2 * (size_t)np) >> l); }
TMX_HD size_t inner_off(uint32_t np, uint32_t l) { return (size_t)np - ((size_t)np >> (l - 1)); }
TMX_HD uint32_t n_header_proofs(uint32_t kind) { return kind == TMX_KIND_SKIP ? 4 : 5; }
TMX_HD size_t sha256_used_chunks(uint32_t kind, uint32_t n_max, uint32_t np) {
    return n_sets(kind) * set_chunks(n_max, np) + (kind == TMX_KIND_SKIP ? 36 : 46);
}

// ---- leaf job: returns false if slot i >= n_max (no SHA-256 call, zero node)
TMX_HD bool sha256_leaf_prepare(const WitnessArgs& a, uint32_t s, uint32_t i, Sha256Hist* hs) {
    uint8_t* node = a.nodes + ((size_t)s * 2 * a.np + i) * 32;
    uint8_t* en = a.node_en + (size_t)s * 2 * a.np + i;
    if (i >= a.n_max) {
        for (int k = 0; k < 32; k++) node[k] = 0;
        *en = 0;
        return false;
    }
    const tmx_offchain_head* h = blob_head(a.blob);
    const bool trusted = (a.kind == TMX_KIND_SKIP && s == 0);
    const uint8_t* pk;
    uint64_t power;
    uint32_t blen, nb;
    if (trusted) {
        const tmx_hash_field* f = blob_hash_fields(a.blob, a.n_max) + i;
        pk = f->pubkey; power = f->voting_power; blen = f->validator_byte_length; nb = h->nb_trusted;
    } else {
        const tmx_validator* v = blob_validators(a.blob) + i;
        pk = v->pubkey; power = v->voting_power; blen = v->validator_byte_length; nb = h->nb_validators;
    }
    uint8_t msg[47], buf[64];
    validator_leaf_message(pk, power, msg);
    int len = 1 + (int)blen;
    if (len > 47) len = 47;
    sha256_pad_blocks(msg, len, buf);
    uint32_t cv[8], st[8];
    for (int k = 0; k < 8; k++) cv[k] = iv256(k);
    sha256_compress_hist(cv, buf, hs, st);
    sha256_state_to_bytes(st, node);
    *en = i < nb ? 1 : 0;
    if (a.log_np == 0)  // single-slot set: the leaf is the root
        for (int k = 0; k < 32; k++) a.aux[AUX_SET_ROOT + 32 * s + k] = node[k];
    return true;
}
TMX_HD size_t sha256_leaf_row0(const WitnessArgs& a, uint32_t s, uint32_t i) {
    return ((size_t)s * set_chunks(a.n_max, a.np) + i) * S256_ROUNDS;
}

// ---- inner node job (level l >= 1, index i): two chunks
TMX_HD void sha256_inner_prepare(const WitnessArgs& a, uint32_t s, uint32_t l, uint32_t i, Sha256Hist hs[2]) {
    uint8_t* base = a.nodes + (size_t)s * 2 * a.np * 32;
    uint8_t* enb = a.node_en + (size_t)s * 2 * a.np;
    const uint8_t* L = base + (level_off(a.np, l - 1) + 2 * i) * 32;
    const uint8_t* R = L + 32;
    const uint8_t enL = enb[level_off(a.np, l - 1) + 2 * i], enR = enb[level_off(a.np, l - 1) + 2 * i + 1];
    uint8_t msg[65], buf[128];
    msg[0] = 1;
    for (int k = 0; k < 32; k++) { msg[1 + k] = L[k]; msg[33 + k] = R[k]; }
    sha256_pad_blocks(msg, 65, buf);
    uint32_t cv[8], st[8];
    for (int k = 0; k < 8; k++) cv[k] = iv256(k);
    sha256_compress_hist(cv, buf, &hs[0], st);
    sha256_compress_hist(st, buf + 64, &hs[1], st);
    uint8_t* out = base + (level_off(a.np, l) + i) * 32;
    if (enL && enR)
        sha256_state_to_bytes(st, out);
    else
        for (int k = 0; k < 32; k++) out[k] = L[k];
    enb[level_off(a.np, l) + i] = enL;
    if (l == a.log_np) {
        uint8_t* root = a.aux + AUX_SET_ROOT + 32 * s;
        for (int k = 0; k < 32; k++) root[k] = out[k];
    }
}
TMX_HD size_t sha256_inner_row0(const WitnessArgs& a, uint32_t s, uint32_t l, uint32_t i) {
    return ((size_t)s * set_chunks(a.n_max, a.np) + a.n_max + 2 * (inner_off(a.np, l) + i)) * S256_ROUNDS;
}

// ---- header proof k: 5 messages (leaf + 4 inner); message j is prepared from the digest of message j - 1
struct HeaderProofDesc {
    uint8_t leaf_msg[80];
    int leaf_len;
    const uint8_t (*aunts)[32];
    unsigned index;
    size_t chunk0;
};
TMX_HD void header_proof_desc(const WitnessArgs& a, uint32_t k, HeaderProofDesc* d) {
    const tmx_offchain_head* h = blob_head(a.blob);
    size_t chunk = n_sets(a.kind) * set_chunks(a.n_max, a.np);
    // which logical proof: 0 aux-valhash(skip) 1 valhash 2 chain 3 height 4 last-block-id 5 aux-next-valhash(step)
    int which;
    if (a.kind == TMX_KIND_SKIP) {
        const int order[4] = {0, 1, 2, 3};
        which = order[k];
        chunk += 9 * (size_t)k;
    } else {
        const int order[5] = {1, 2, 3, 4, 5};
        which = order[k];
        chunk += 9 * (size_t)k + (k > 3 ? 1 : 0);
    }
    d->chunk0 = chunk;
    for (int i = 0; i < 80; i++) d->leaf_msg[i] = 0;
    const tmx_hash_proof* hp = nullptr;
    switch (which) {
        case 0: hp = &h->aux_hash_proof; d->index = TMX_VALIDATORS_HASH_INDEX; break;
        case 1: hp = &h->validators_hash_proof; d->index = TMX_VALIDATORS_HASH_INDEX; break;
        case 5: hp = &h->aux_hash_proof; d->index = TMX_NEXT_VALIDATORS_HASH_INDEX; break;
        default: break;
    }
    if (hp) {
        for (int i = 0; i < 34; i++) d->leaf_msg[1 + i] = hp->leaf[i];
        d->leaf_len = 35;
        d->aunts = hp->aunts;
    } else if (which == 2) {
        for (int i = 0; i < 52; i++) d->leaf_msg[1 + i] = h->chain_id_proof.chain_id[i];
        int len = (int)h->chain_id_proof.enc_chain_id_byte_length + 1;
        d->leaf_len = len > 55 ? 55 : len;
        d->aunts = h->chain_id_proof.aunts;
        d->index = TMX_CHAIN_ID_INDEX;
    } else if (which == 3) {
        d->leaf_msg[1] = 0x08;
        marshal_int64_varint(h->height_proof.height, d->leaf_msg + 2);
        int len = (int)h->height_proof.enc_height_byte_length + 1;
        d->leaf_len = len > 55 ? 55 : len;
        d->aunts = h->height_proof.aunts;
        d->index = TMX_BLOCK_HEIGHT_INDEX;
    } else {
        for (int i = 0; i < 72; i++) d->leaf_msg[1 + i] = h->last_block_id_proof.leaf[i];
        d->leaf_len = 73;
        d->aunts = h->last_block_id_proof.aunts;
        d->index = TMX_LAST_BLOCK_ID_INDEX;
    }
}
// message j of the proof (j = 0 leaf, 1..4 inner).  cur: running digest (in/out).  Returns the chunk count.
TMX_HD int header_proof_prepare(const HeaderProofDesc& d, int j, uint8_t cur[32], Sha256Hist hs[2]) {
    uint8_t msg[80], buf[128];
    int len;
    if (j == 0) {
        len = d.leaf_len;
        for (int i = 0; i < len; i++) msg[i] = d.leaf_msg[i];
    } else {
        const uint8_t* aunt = d.aunts[j - 1];
        const bool right = (d.index >> (j - 1)) & 1;  // bit set: node is the right child, aunt on the left
        msg[0] = 1;
        for (int i = 0; i < 32; i++) {
            msg[1 + i] = right ? aunt[i] : cur[i];
            msg[33 + i] = right ? cur[i] : aunt[i];
        }
        len = 65;
    }
    const int nb = sha256_pad_blocks(msg, len, buf);
    uint32_t st[8];
    for (int k = 0; k < 8; k++) st[k] = iv256(k);
    for (int b = 0; b < nb; b++) sha256_compress_hist(st, buf + 64 * b, &hs[b], st);
    sha256_state_to_bytes(st, cur);
    return nb;
}

TMX_HD void sha256_padding_prepare(Sha256Hist* hs) {
    uint8_t zero[64];
    for (int i = 0; i < 64; i++) zero[i] = 0;
    uint32_t cv[8], st[8];
    for (int k = 0; k < 8; k++) cv[k] = iv256(k);
    sha256_compress_hist(cv, zero, hs, st);
}
TMX_HD void sha512_padding_prepare(Sha512Hist* hs) {
    uint8_t zero[128];
    for (int i = 0; i < 128; i++) zero[i] = 0;
    uint64_t cv[8], st[8];
    for (int k = 0; k < 8; k++) cv[k] = iv512(k);
    sha512_compress_hist(cv, zero, hs, st);
    hs->two = 0;
}

// ---- validator slot i: the triple the Ed25519 gadget verifies (REF conversion.rs:79-134: unsigned slots carry
// the dummy signature; the gadget swaps in the dummy key and 32-byte zero message for them)
struct EdTriple {
    uint8_t pk[32], sig[64], msg[TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX];
    int len;
};
TMX_HD void effective_triple(const tmx_validator* v, EdTriple* t) {
    if (v->is_signed) {
        for (int i = 0; i < 32; i++) { t->pk[i] = v->pubkey[i]; t->sig[i] = v->sig_r[i]; t->sig[32 + i] = v->sig_s[i]; }
        for (int i = 0; i < TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX; i++) t->msg[i] = v->message[i];
        t->len = v->message_byte_length > TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX ? TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX
                                                                                   : (int)v->message_byte_length;
    } else {
        for (int i = 0; i < 32; i++) t->pk[i] = dummy_pk(i);
        for (int i = 0; i < 64; i++) t->sig[i] = dummy_sig(i);
        for (int i = 0; i < TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX; i++) t->msg[i] = 0;
        t->len = 32;
    }
}

// SHA-512 of R || A || M into two chunk histories (second one is the zero-block filler when the message
// fits one block); digest = big-endian state after the real blocks.
TMX_HD void sha512_validator_prepare(const EdTriple& t, Sha512Hist hs[2], uint8_t digest[64]) {
    uint8_t m[64 + TMX_VALIDATOR_MESSAGE_BYTES_LENGTH_MAX], buf[256];
    for (int i = 0; i < 32; i++) { m[i] = t.sig[i]; m[32 + i] = t.pk[i]; }
    for (int i = 0; i < t.len; i++) m[64 + i] = t.msg[i];
    const int nb = sha512_pad_blocks(m, 64 + t.len, buf);
    uint64_t st[8];
    for (int k = 0; k < 8; k++) st[k] = iv512(k);
    sha512_compress_hist(st, buf, &hs[0], st);
    if (nb == 2)
        sha512_compress_hist(st, buf + 128, &hs[1], st);
    else
        sha512_padding_prepare(&hs[1]);
    hs[0].two = hs[1].two = (nb == 2);
    for (int k = 0; k < 8; k++)
        for (int j = 0; j < 8; j++) digest[8 * k + j] = (uint8_t)(st[k] >> (56 - 8 * j));
}

// What the sequential phase leaves behind for one validator slot: scalars, addend cells, the points the logic table needs
// and the verdict of the signature equation [s]B + [h](-A) == R (cofactorless, canonical encodings required).
struct EdSlotInfo {
    uint64_t s[4], h[4];
    EdAddendTable tab;
    fe256 xA, yA, xR, yR, xD, yD;
    fe256 QX, QY, QZ;
    uint8_t digest[64];
    uint32_t ok;
    uint32_t pad;
};
// everything but the 256 sequential rows: effective triple -> h, s, decompressed A and R, addend table
TMX_HD void ed_slot_prepare(const EdTriple& t, const uint8_t digest[64], EdSlotInfo* e, ge_cached51 tab[4], ge51* R_out) {
    fe256 sb = fe256_from_bytes(t.sig + 32);
    for (int i = 0; i < 4; i++) e->s[i] = sb.w[i];
    sc_reduce512(digest, e->h);
    for (int i = 0; i < 64; i++) e->digest[i] = digest[i];
    bool ok = sc_lt_l(e->s);
    ge51 A, R;
    if (!ge_decompress51(t.pk, &A)) {
        ok = false;
        A = ge_identity51();
    }
    if (!ge_decompress51(t.sig, &R)) {
        ok = false;
        R = ge_identity51();
    }
    e->xA = fe_freeze(A.X); e->yA = fe_freeze(A.Y);
    e->xR = fe_freeze(R.X); e->yR = fe_freeze(R.Y);
    ed_slot_table(A, tab, &e->tab, &e->xD, &e->yD);
    e->ok = ok ? 1u : 0u;
    *R_out = R;
}
// Q = (X : Y : Z) equals the affine R?
TMX_HD bool ed_result_equals(const ge_acc51& q, const ge51& R) {
    return fe256_eq(fe_freeze(fe_mul(R.X, q.Z)), fe_freeze(q.X)) && fe256_eq(fe_freeze(fe_mul(R.Y, q.Z)), fe_freeze(q.Y));
}

}  // namespace tmx

// host entry points of witness.cu used by the circuit driver (phase-split so the latency-bound Ed25519 ladders can
// run on a second stream while the SHA-256 table is being proved)
struct tmx_ctx;
namespace tmx {
#if defined(__CUDACC__)
size_t witness_points_bytes(uint32_t n_max);
int witness_make_args(tmx_ctx* ctx, const uint8_t* d_blob, uint32_t kind, uint32_t n_max, uint64_t* t256, uint64_t* t512,
                      uint64_t* ted, uint8_t* d_aux, WitnessArgs* a);
int witness_run_sha256(tmx_ctx* ctx, const WitnessArgs& a, cudaStream_t st);
int run_ed25519_ladder(tmx_ctx* ctx, const WitnessArgs& a, void* points, cudaStream_t st);
int run_ed25519_expand(tmx_ctx* ctx, const WitnessArgs& a, const void* points, cudaStream_t st);
// layout of the `points` scratch: EdSlotInfo[n_slots], then ge_acc_packed[n_slots][256]
size_t witness_slot_count(uint32_t n_max);
#endif
}  // namespace tmx
