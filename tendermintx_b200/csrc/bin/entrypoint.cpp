// `skip` / `step` entry points: the reference's bin/skip.rs and bin/step.rs
// [REF bin/skip.rs:22-27, bin/step.rs:21-26: SkipCircuit::<VALIDATOR_SET_SIZE_MAX, .., CelestiaConfig>::entrypoint()]
// with plonky2x's two subcommands [REF succinct.json:8-9,15-16]:
//
//     ./build/skip build                 -> ./build/main.circuit
//     ./build/skip prove input.json      -> output.json
//
// input.json  = {"type": "req_bytes", "data": {"input": "0x<48 bytes (skip) | 40 bytes (step)>"}}
// output.json = {"type": "res_bytes", "data": {"proof": "0x...", "output": "0x<32 bytes>"}}
//
// A native host over the C ABI only (include/tmx.h): no Python, no torch.  The circuit kind is fixed at compile time
// (-DTMX_ENTRY_KIND=0 step / 1 skip), the generics of the reference are flags with the reference's defaults:
// --n-max (VALIDATOR_SET_SIZE_MAX = 100 [REF circuits/consts.rs:4]), --chain-id (celestia), --skip-max (100800
// [REF circuits/config.rs:12]).  The off-chain inputs come from a fixture directory in the RPC JSON layout
// (--fixtures or TMX_FIXTURE_DIR): the reference's RPC mode needs a network.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sys/stat.h>
#include "../../../include/tmx.h"

#ifndef TMX_ENTRY_KIND
#define TMX_ENTRY_KIND 1
#endif

static int die(const char* what) {
    fprintf(stderr, "error: %s: %s\n", what, tmx_last_error());
    return 1;
}

static std::string hex(const uint8_t* p, size_t n) {
    static const char* d = "0123456789abcdef";
    std::string s = "0x";
    for (size_t i = 0; i < n; i++) {
        s += d[p[i] >> 4];
        s += d[p[i] & 15];
    }
    return s;
}

// the value of "input" inside {"type": "req_bytes", "data": {"input": "0x.."}}
static bool read_input(const char* path, std::vector<uint8_t>* out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    std::string s;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    if (s.find("\"req_bytes\"") == std::string::npos) return false;
    size_t k = s.find("\"input\"");
    if (k == std::string::npos) return false;
    k = s.find('"', s.find(':', k));
    if (k == std::string::npos) return false;
    size_t e = s.find('"', k + 1);
    if (e == std::string::npos) return false;
    std::string h = s.substr(k + 1, e - k - 1);
    if (h.rfind("0x", 0) == 0) h = h.substr(2);
    if (h.size() % 2) return false;
    out->clear();
    for (size_t i = 0; i < h.size(); i += 2) {
        unsigned v;
        if (sscanf(h.c_str() + i, "%2x", &v) != 1) return false;
        out->push_back((uint8_t)v);
    }
    return true;
}

int main(int argc, char** argv) {
    const uint32_t kind = TMX_ENTRY_KIND;
    uint32_t n_max = 100;
    std::string chain_id = "celestia", circuit_path = "build/main.circuit", out_path = "output.json", input_path, cmd;
    const char* env_fix = getenv("TMX_FIXTURE_DIR");
    std::string fixtures = env_fix ? env_fix : "";
    uint64_t skip_max = 100800;
    int device = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> const char* {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: %s needs a value\n", name);
                exit(2);
            }
            return argv[++i];
        };
        if (a == "--n-max") n_max = (uint32_t)atoi(val("--n-max"));
        else if (a == "--chain-id") chain_id = val("--chain-id");
        else if (a == "--skip-max") skip_max = strtoull(val("--skip-max"), nullptr, 10);
        else if (a == "--fixtures") fixtures = val("--fixtures");
        else if (a == "--circuit-file") circuit_path = val("--circuit-file");
        else if (a == "--out") out_path = val("--out");
        else if (a == "--device") device = atoi(val("--device"));
        else if (a == "--input-json") input_path = val("--input-json");
        else if (cmd.empty()) cmd = a;
        else if (input_path.empty()) input_path = a;
    }
    if (cmd != "build" && cmd != "prove") {
        fprintf(stderr, "usage: %s build | prove input.json  [--n-max N] [--chain-id ID] [--fixtures DIR] [--circuit-file PATH] [--out PATH]\n", argv[0]);
        return 2;
    }
    tmx_ctx* ctx = nullptr;
    if (tmx_ctx_create(device, &ctx)) return die("tmx_ctx_create");
    tmx_circuit* c = nullptr;
    if (cmd == "build") {
        if (tmx_circuit_build(ctx, kind, n_max, chain_id.c_str(), chain_id.size(), skip_max, &c)) return die("build");
        const size_t slash = circuit_path.rfind('/');
        if (slash != std::string::npos) mkdir(circuit_path.substr(0, slash).c_str(), 0755);
        if (tmx_circuit_save(c, circuit_path.c_str())) return die("save");
        uint64_t dg[4];
        tmx_circuit_digest(c, dg);
        printf("wrote %s (%s circuit, VALIDATOR_SET_SIZE_MAX = %u, chain id %s, digest %016llx%016llx%016llx%016llx)\n", circuit_path.c_str(),
               kind ? "skip" : "step", n_max, chain_id.c_str(), (unsigned long long)dg[0], (unsigned long long)dg[1],
               (unsigned long long)dg[2], (unsigned long long)dg[3]);
    } else {
        if (input_path.empty() || fixtures.empty()) {
            fprintf(stderr, "error: prove needs input.json and --fixtures DIR (or TMX_FIXTURE_DIR)\n");
            return 2;
        }
        std::vector<uint8_t> input;
        if (!read_input(input_path.c_str(), &input)) {
            fprintf(stderr, "error: %s is not a req_bytes request\n", input_path.c_str());
            return 1;
        }
        if (tmx_circuit_load(ctx, circuit_path.c_str(), &c)) return die("load");
        tmx_proof* proof = nullptr;
        uint8_t out[32];
        if (tmx_prove_fixture(c, input.data(), input.size(), fixtures.c_str(), &proof, out)) {
            fprintf(stderr, "error: prove: %s (failing check id %d)\n", tmx_last_error(), tmx_last_check());
            return 1;
        }
        std::vector<uint8_t> bytes(tmx_proof_size(proof));
        if (tmx_proof_bytes(proof, bytes.data(), bytes.size())) return die("proof bytes");
        tmx_proof_free(proof);
        if (tmx_verify(c, bytes.data(), bytes.size(), input.data(), input.size(), out)) return die("verify");
        FILE* f = fopen(out_path.c_str(), "wb");
        if (!f) {
            fprintf(stderr, "error: cannot write %s\n", out_path.c_str());
            return 1;
        }
        fprintf(f, "{\"type\": \"res_bytes\", \"data\": {\"proof\": \"%s\", \"output\": \"%s\"}}\n", hex(bytes.data(), bytes.size()).c_str(),
                hex(out, 32).c_str());
        fclose(f);
        printf("wrote %s: output %s (%zu proof bytes)\n", out_path.c_str(), hex(out, 32).c_str(), bytes.size());
    }
    tmx_circuit_free(c);
    tmx_ctx_destroy(ctx);
    return 0;
}
