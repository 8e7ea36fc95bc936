// Circuit definition as data (see circuit_def.cuh): symbolic run of the AIR templates, constant columns and their
// commitment, digest, artefact (de)serialisation, and a host interpreter of the DAG.  Host code only.
#include "circuit_def.cuh"
#include "params.cuh"
#include "poseidon.cuh"
#include "witness.cuh"
#include "logic.cuh"
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace tmx {

void set_error(const std::string& msg);

// ------------------------------------------------------------------------------------------ symbolic field
struct SymBuilder {
    std::vector<SymNode> nodes;
    std::unordered_map<uint64_t, std::vector<uint32_t>> index;  // hash of (op, a, b, val) -> candidates
    std::vector<uint64_t> prog;
    uint32_t n_constraints = 0, n_helpers = 0;
    std::string error;

    uint32_t intern(uint32_t op, uint32_t a, uint32_t b, uint64_t val, uint32_t deg) {
        const uint64_t h = ((uint64_t)op * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)a << 32 | b) ^ (val * 0xC2B2AE3D27D4EB4FULL);
        auto& cand = index[h];
        for (uint32_t id : cand) {
            const SymNode& n = nodes[id];
            if (n.op == op && n.a == a && n.b == b && n.val == val) return id;
        }
        nodes.push_back(SymNode{op, a, b, deg, val});
        cand.push_back((uint32_t)nodes.size() - 1);
        return (uint32_t)nodes.size() - 1;
    }
    uint32_t constant(uint64_t v) { return intern(SYM_CONST, 0, 0, v % GL_P, 0); }
    uint32_t col(uint32_t src, uint32_t c) { return intern(SYM_COL, src, c, 0, 1); }
    bool is_const(uint32_t id, uint64_t* v) const {
        if (nodes[id].op != SYM_CONST) return false;
        *v = nodes[id].val;
        return true;
    }
    uint32_t add(uint32_t a, uint32_t b) {
        uint64_t x, y;
        const bool ca = is_const(a, &x), cb = is_const(b, &y);
        if (ca && cb) return constant(gl_add(x, y));
        if (ca && x == 0) return b;
        if (cb && y == 0) return a;
        if (a > b) std::swap(a, b);  // commutative: canonical operand order
        return intern(SYM_ADD, a, b, 0, std::max(nodes[a].deg, nodes[b].deg));
    }
    uint32_t sub(uint32_t a, uint32_t b) {
        uint64_t x, y;
        const bool ca = is_const(a, &x), cb = is_const(b, &y);
        if (ca && cb) return constant(gl_sub(x, y));
        if (cb && y == 0) return a;
        if (a == b) return constant(0);
        return intern(SYM_SUB, a, b, 0, std::max(nodes[a].deg, nodes[b].deg));
    }
    uint32_t mul(uint32_t a, uint32_t b) {
        uint64_t x, y;
        const bool ca = is_const(a, &x), cb = is_const(b, &y);
        if (ca && cb) return constant(gl_mul(x, y));
        if ((ca && x == 0) || (cb && y == 0)) return constant(0);
        if (ca && x == 1) return b;
        if (cb && y == 1) return a;
        if (a > b) std::swap(a, b);
        return intern(SYM_MUL, a, b, 0, nodes[a].deg + nodes[b].deg);
    }
    void fail(const std::string& m) {
        if (error.empty()) error = m;
    }
};
static thread_local SymBuilder* g_sym = nullptr;

struct Sym {
    uint32_t id;
    static Sym mk(uint32_t i) { Sym s; s.id = i; return s; }
    static Sym c(uint64_t x) { return mk(g_sym->constant(x)); }
};
static inline Sym operator+(Sym a, Sym b) { return Sym::mk(g_sym->add(a.id, b.id)); }
static inline Sym operator-(Sym a, Sym b) { return Sym::mk(g_sym->sub(a.id, b.id)); }
static inline Sym operator*(Sym a, Sym b) { return Sym::mk(g_sym->mul(a.id, b.id)); }

struct SymRow {
    uint32_t src;
    Sym operator[](int c) const { return Sym::mk(g_sym->col(src, (uint32_t)c)); }
};
struct SymEmit {
    void operator()(Sym c) const {
        if (g_sym->nodes[c.id].deg > 3) g_sym->fail("constraint " + std::to_string(g_sym->n_constraints) + " has degree > 3");
        g_sym->prog.push_back(0);
        g_sym->prog.push_back(c.id);
        g_sym->n_constraints++;
    }
};
struct SymBus {
    template <class Tup>
    static uint32_t push_tuple(Sym tag, Sym m, int len, const Tup& tup) {
        if (g_sym->nodes[tag.id].deg > 1) g_sym->fail("bus tag of degree > 1");
        g_sym->prog.push_back(tag.id);
        g_sym->prog.push_back(m.id);
        g_sym->prog.push_back((uint64_t)len);
        uint32_t deg = 0;
        for (int i = 0; i < len; i++) {
            const Sym v = tup(i);
            deg = std::max(deg, g_sym->nodes[v.id].deg);
            g_sym->prog.push_back(v.id);
        }
        return deg;
    }
    template <class Tup>
    void one(Sym tag, Sym m, int len, const Tup& tup) {
        g_sym->prog.push_back(1);
        const uint32_t df = std::max(push_tuple(tag, m, len, tup), g_sym->nodes[tag.id].deg);
        if (1 + df > 3 || g_sym->nodes[m.id].deg > 3) g_sym->fail("bus.one: degree > 3 (helper " + std::to_string(g_sym->n_helpers) + ")");
        g_sym->n_helpers++;
    }
    template <class TA, class TB>
    void two(Sym tag_a, Sym ma, int len_a, const TA& ta, Sym tag_b, Sym mb, int len_b, const TB& tb) {
        g_sym->prog.push_back(2);
        const uint32_t da = std::max(push_tuple(tag_a, ma, len_a, ta), g_sym->nodes[tag_a.id].deg);
        const uint32_t db = std::max(push_tuple(tag_b, mb, len_b, tb), g_sym->nodes[tag_b.id].deg);
        const uint32_t dma = g_sym->nodes[ma.id].deg, dmb = g_sym->nodes[mb.id].deg;
        if (1 + da + db > 3 || dma + db > 3 || dmb + da > 3)
            g_sym->fail("bus.two: degree > 3 (helper " + std::to_string(g_sym->n_helpers) + ")");
        g_sym->n_helpers++;
    }
    void two_lookups(Sym tag_a, Sym va, Sym tag_b, Sym vb) {
        const Sym m = Sym::c(GL_P - 1);
        two(tag_a, m, 1, [&](int) { return va; }, tag_b, m, 1, [&](int) { return vb; });
    }
};

// ------------------------------------------------------------------------------------------ host commitment of constant columns
template <class Fn>
static void parallel_for(size_t n, const Fn& fn) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    if (nt > 16) nt = 16;
    if (n < 4096 || nt == 1) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([&fn, lo, hi] { fn(lo, hi); });
    }
    for (auto& x : th) x.join();
}

// LDE (rate 1/2, coset 7, bit-reversed row order) of n_cols columns and the Merkle cap of its rows: the host twin of
// tmx_lde + tmx_poseidon_merkle for the (small) constant-column batches
static std::vector<gl> host_commit_columns(const std::vector<gl>& cols, size_t n_cols, unsigned log_n) {
    const size_t n = (size_t)1 << log_n, m = n << STARK_RATE_BITS;
    const unsigned km = log_n + STARK_RATE_BITS;
    std::vector<gl> lde(n_cols * m);
    for (size_t c = 0; c < n_cols; c++) {
        std::vector<gl> a(cols.begin() + c * n, cols.begin() + (c + 1) * n);
        air_host_ntt(a, true);
        gl s = 1;
        for (size_t i = 0; i < n; i++) {
            a[i] = gl_mul(a[i], s);
            s = gl_mul(s, GL_GEN);
        }
        a.resize(m, 0);
        air_host_ntt(a, false);
        for (size_t j = 0; j < m; j++) lde[c * m + j] = a[bitrev32((uint32_t)j, km)];
    }
    const unsigned cap_h = std::min<unsigned>(km, STARK_CAP_HEIGHT);
    std::vector<gl> level(4 * m);
    parallel_for(m, [&](size_t lo, size_t hi) {
        for (size_t j = lo; j < hi; j++) poseidon_hash_row(lde.data() + j, m, n_cols, &level[4 * j]);
    });
    size_t rows = m;
    while (rows > ((size_t)1 << cap_h)) {
        std::vector<gl> up(4 * (rows / 2));
        parallel_for(rows / 2, [&](size_t lo, size_t hi) {
            for (size_t j = lo; j < hi; j++) poseidon_two_to_one(&level[8 * j], &level[8 * j + 4], &up[4 * j]);
        });
        level.swap(up);
        rows /= 2;
    }
    return level;
}

// ------------------------------------------------------------------------------------------ building
static void hash_words(const std::vector<gl>& in, gl out[4]) {
    gl s[12] = {0};
    for (size_t off = 0; off < in.size(); off += 8) {
        const size_t k = std::min<size_t>(8, in.size() - off);
        for (size_t i = 0; i < k; i++) s[i] = in[off + i] % GL_P;
        poseidon_permute(s);
    }
    for (int i = 0; i < 4; i++) out[i] = s[i];
}

static void header_words(const CircuitDef& c, std::vector<uint64_t>& w) {
    w.push_back(STARK_CIRCUIT_MAGIC);
    w.push_back(c.kind);
    w.push_back(c.n_max);
    w.push_back(c.skip_max);
    w.push_back(c.chain_id.size());
    for (int i = 0; i < 8; i++) {
        uint64_t x = 0;
        for (int j = 0; j < 8; j++) {
            const size_t k = 8 * i + j;
            if (k < c.chain_id.size()) x |= (uint64_t)(uint8_t)c.chain_id[k] << (8 * j);
        }
        w.push_back(x);
    }
    w.push_back(STARK_RATE_BITS);
    w.push_back(STARK_CAP_HEIGHT);
    w.push_back(STARK_POW_BITS);
    w.push_back(STARK_NUM_QUERIES);
    w.push_back(STARK_ARITY_BITS);
    w.push_back(STARK_FINAL_POLY_BITS);
    w.push_back(TMX_N_TABLES);
}
static void table_words(const TableDef& t, bool with_constants, std::vector<uint64_t>& w) {
    w.push_back(t.n_main ? 1 : 0);
    if (!t.n_main) return;
    const uint64_t f[] = {t.log_n, t.n_main, t.n_const, t.n_per, t.period, t.n_helpers, t.n_constraints,
                          t.const_cap.size(), t.nodes.size(), t.prog.size()};
    w.insert(w.end(), f, f + 10);
    w.insert(w.end(), t.periodic.begin(), t.periodic.end());
    if (with_constants) w.insert(w.end(), t.constants.begin(), t.constants.end());
    w.insert(w.end(), t.const_cap.begin(), t.const_cap.end());
    for (const SymNode& n : t.nodes) {
        w.push_back((uint64_t)n.op | ((uint64_t)n.deg << 8) | ((uint64_t)n.a << 16));
        w.push_back(n.op == SYM_CONST ? n.val : (uint64_t)n.b);
    }
    w.insert(w.end(), t.prog.begin(), t.prog.end());
}
static void compute_digest(CircuitDef& c) {
    std::vector<uint64_t> w;
    header_words(c, w);
    for (int t = 0; t < TMX_N_TABLES; t++) table_words(c.tables[t], false, w);
    hash_words(w, c.digest);
}

std::vector<uint64_t> CircuitDef::serialize() const {
    std::vector<uint64_t> w;
    header_words(*this, w);
    for (int t = 0; t < TMX_N_TABLES; t++) table_words(tables[t], true, w);
    for (int i = 0; i < 4; i++) w.push_back(digest[i]);
    return w;
}

static bool build_table(int table, AirShape sh, TableDef& t, std::string& err) {
    const size_t n = air_table_rows(table, sh);
    if (!n) return true;
    t.log_n = ilog2(n);
    t.n_main = (uint32_t)air_table_cols(table, sh);
    t.n_const = (uint32_t)air_table_const_cols(table, sh);
    t.n_per = (uint32_t)air_n_periodic(table);
    t.period = (uint32_t)air_period(table);
    // symbolic run
    SymBuilder b;
    g_sym = &b;
    SymRow l{SRC_LOCAL}, nx{SRC_NEXT}, k{SRC_CONST}, per{SRC_PERIODIC};
    SymEmit emit;
    SymBus bus;
    air_eval_any<Sym>(table, sh, l, nx, k, per, emit, bus);
    g_sym = nullptr;
    if (!b.error.empty()) {
        err = "table " + std::to_string(table) + ": " + b.error;
        return false;
    }
    t.n_constraints = b.n_constraints;
    t.n_helpers = b.n_helpers;
    if ((int)t.n_helpers != air_table_helpers(table, sh)) {
        err = "table " + std::to_string(table) + ": helper count " + std::to_string(t.n_helpers) + " differs from air_table_helpers";
        return false;
    }
    t.nodes.swap(b.nodes);
    t.prog.swap(b.prog);
    t.periodic.resize((size_t)t.n_per * t.period);
    for (uint32_t pc = 0; pc < t.n_per; pc++)
        for (uint32_t r = 0; r < t.period; r++) t.periodic[(size_t)pc * t.period + r] = air_periodic_pattern(table, (int)pc, r, h_K256, h_K512);
    t.constants.resize((size_t)t.n_const * n);
    parallel_for(n, [&](size_t lo, size_t hi) {
        for (uint32_t kc = 0; kc < t.n_const; kc++)
            for (size_t r = lo; r < hi; r++) t.constants[(size_t)kc * n + r] = air_table_const_value(table, (int)kc, r, sh);
    });
    if (t.n_const) t.const_cap = host_commit_columns(t.constants, t.n_const, t.log_n);
    return true;
}

std::shared_ptr<const CircuitDef> circuit_def_get(uint32_t kind, uint32_t n_max, const std::string& chain_id, uint64_t skip_max) {
    static std::mutex mu;
    static std::map<std::string, std::shared_ptr<const CircuitDef>> cache;
    const std::string key = std::to_string(kind) + "/" + std::to_string(n_max) + "/" + std::to_string(skip_max) + "/" + chain_id;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    poseidon_generate_constants();
    auto c = std::make_shared<CircuitDef>();
    c->kind = kind;
    c->n_max = n_max;
    c->skip_max = skip_max;
    c->chain_id = chain_id;
    const AirShape sh = air_shape(kind, n_max, chain_id.data(), chain_id.size());
    for (int t = 0; t < TMX_N_TABLES; t++) {
        std::string err;
        if (!build_table(t, sh, c->tables[t], err)) {
            set_error("circuit definition: " + err);
            return nullptr;
        }
    }
    compute_digest(*c);
    cache[key] = c;
    return c;
}

std::shared_ptr<const CircuitDef> circuit_def_parse(const uint64_t* w, size_t nw) {
    size_t pos = 0;
    bool bad = false;
    auto get = [&]() -> uint64_t {
        if (pos >= nw) { bad = true; return 0; }
        return w[pos++];
    };
    auto c = std::make_shared<CircuitDef>();
    if (get() != STARK_CIRCUIT_MAGIC) { set_error("circuit artefact: bad magic"); return nullptr; }
    c->kind = (uint32_t)get();
    c->n_max = (uint32_t)get();
    c->skip_max = get();
    const size_t cl = (size_t)get();
    uint64_t cw[8];
    for (int i = 0; i < 8; i++) cw[i] = get();
    if (cl > 50 || bad) { set_error("circuit artefact: bad header"); return nullptr; }
    for (size_t k = 0; k < cl; k++) c->chain_id.push_back((char)(uint8_t)(cw[k / 8] >> (8 * (k % 8))));
    const uint64_t want[] = {STARK_RATE_BITS, STARK_CAP_HEIGHT, STARK_POW_BITS, (uint64_t)STARK_NUM_QUERIES, STARK_ARITY_BITS,
                             STARK_FINAL_POLY_BITS, TMX_N_TABLES};
    for (uint64_t x : want)
        if (get() != x) { set_error("circuit artefact: protocol parameters differ from this build"); return nullptr; }
    for (int ti = 0; ti < TMX_N_TABLES && !bad; ti++) {
        TableDef& t = c->tables[ti];
        if (!get()) continue;
        uint64_t f[10];
        for (int i = 0; i < 10; i++) f[i] = get();
        if (bad || f[0] > 28 || f[1] > (1u << 20) || f[2] > 4096 || f[3] > 64 || f[4] > (1u << 20) || f[7] > 4096 || f[8] > (1u << 26) || f[9] > (1u << 28)) {
            bad = true;
            break;
        }
        t.log_n = (uint32_t)f[0]; t.n_main = (uint32_t)f[1]; t.n_const = (uint32_t)f[2]; t.n_per = (uint32_t)f[3];
        t.period = (uint32_t)f[4]; t.n_helpers = (uint32_t)f[5]; t.n_constraints = (uint32_t)f[6];
        auto take = [&](std::vector<gl>& dst, size_t cnt) {
            if (pos + cnt > nw) { bad = true; return; }
            dst.assign(w + pos, w + pos + cnt);
            pos += cnt;
        };
        take(t.periodic, (size_t)t.n_per * t.period);
        take(t.constants, (size_t)t.n_const << t.log_n);
        take(t.const_cap, (size_t)f[7]);
        if (bad || pos + 2 * f[8] > nw) { bad = true; break; }
        t.nodes.resize((size_t)f[8]);
        for (size_t i = 0; i < t.nodes.size(); i++) {
            const uint64_t a = get(), b = get();
            SymNode& n = t.nodes[i];
            n.op = (uint32_t)(a & 0xFF);
            n.deg = (uint32_t)((a >> 8) & 0xFF);
            n.a = (uint32_t)(a >> 16);
            n.b = n.op == SYM_CONST ? 0 : (uint32_t)b;
            n.val = n.op == SYM_CONST ? b : 0;
            if (n.op > SYM_MUL || (n.op >= SYM_ADD && (n.a >= i || n.b >= i))) bad = true;
        }
        take(t.prog, (size_t)f[9]);
    }
    gl dg[4];
    for (int i = 0; i < 4; i++) dg[i] = get();
    if (bad || pos != nw) { set_error("circuit artefact: truncated or malformed"); return nullptr; }
    poseidon_generate_constants();
    compute_digest(*c);
    if (memcmp(dg, c->digest, sizeof dg)) { set_error("circuit artefact: digest mismatch"); return nullptr; }
    return c;
}

void circuit_def_eval_constraints(const TableDef& t, const gl* local, const gl* next, const gl* consts, const gl* periodic,
                                  std::vector<gl>& out) {
    std::vector<gl> v(t.nodes.size());
    for (size_t i = 0; i < t.nodes.size(); i++) {
        const SymNode& n = t.nodes[i];
        switch (n.op) {
            case SYM_CONST: v[i] = n.val; break;
            case SYM_COL: v[i] = n.a == SRC_LOCAL ? local[n.b] : n.a == SRC_NEXT ? next[n.b] : n.a == SRC_CONST ? consts[n.b] : periodic[n.b]; break;
            case SYM_ADD: v[i] = gl_add(v[n.a], v[n.b]); break;
            case SYM_SUB: v[i] = gl_sub(v[n.a], v[n.b]); break;
            default: v[i] = gl_mul(v[n.a], v[n.b]); break;
        }
    }
    out.clear();
    for (size_t p = 0; p < t.prog.size();) {
        const uint64_t kind = t.prog[p++];
        if (kind == 0) {
            out.push_back(v[t.prog[p++]]);
        } else {
            for (uint64_t part = 0; part < kind; part++) p += 3 + t.prog[p + 2];
        }
    }
}

}  // namespace tmx

using namespace tmx;

// The build artefact of a circuit shape without a GPU (tests, the CPU oracle's input): returns the number of u64 words,
// copies at most `cap` of them.
extern "C" size_t tmx_circuit_artefact(uint32_t kind, uint32_t n_max, const char* chain_id, size_t chain_id_len, uint64_t skip_max,
                                       uint64_t* out, size_t cap) {
    if (!chain_id || kind > 1 || n_max == 0 || n_max > 4096 || chain_id_len == 0 || chain_id_len > 50) return 0;
    auto def = circuit_def_get(kind, n_max, std::string(chain_id, chain_id_len), skip_max);
    if (!def) return 0;
    const std::vector<uint64_t> w = def->serialize();
    if (out) memcpy(out, w.data(), std::min(cap, w.size()) * sizeof(uint64_t));
    return w.size();
}
