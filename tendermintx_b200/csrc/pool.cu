// Prover pool: several independent proofs in flight on ONE GPU.
//
// A proof is a chain of kernels interleaved with Fiat-Shamir round trips to the host (commitment cap -> challenge ->
// next kernel), so a single prover leaves the GPU idle for about a sixth of the proof.  Proofs are independent
// [SURVEY section 8e: each skip / step proof depends only on its own inputs], so a proving service keeps a few of them
// going: every prover has its own tmx_ctx (stream, scratch, tables), its own circuit buffers and its own host thread;
// the gaps of one are filled by the kernels of the others.  Nothing is shared between provers and no proof changes
// (tests/test_gpu_prove.py compares the bytes).  The reference has no counterpart: upstream the same role is played by
// the Succinct platform dispatching `prove` requests to workers [REF bin/tendermintx.rs:169-223, succinct.json].
#include "ctx.cuh"
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

namespace tmx {
void set_last_check(int check);
}

struct tmx_pool {
    struct Job {
        std::vector<uint8_t> input, blob;
        bool resident = false, done = false;
        int rc = 0, check = 0;
        std::string err;
        tmx_proof* proof = nullptr;
        uint8_t out[32];
    };
    std::vector<tmx_ctx*> ctxs;
    std::vector<tmx_circuit*> circuits;
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::deque<std::shared_ptr<Job>> queue;
    std::map<uint64_t, std::shared_ptr<Job>> jobs;
    uint64_t next_ticket = 1;
    unsigned in_progress = 0;  // jobs a worker has taken off the queue and is still proving
    bool stop = false;

    void run(size_t k) {
        for (;;) {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return stop || !queue.empty(); });
                if (queue.empty()) return;  // stop requested and nothing left
                j = queue.front();
                queue.pop_front();
                in_progress++;
            }
            j->rc = tmx_prove(circuits[k], j->input.data(), j->input.size(), j->resident ? nullptr : j->blob.data(),
                              j->resident ? 0 : j->blob.size(), &j->proof, j->out);
            if (j->rc) {
                j->err = tmx_last_error();
                j->check = tmx_last_check();
            }
            {
                std::lock_guard<std::mutex> lk(m);
                j->done = true;
                in_progress--;
            }
            cv_done.notify_all();
        }
    }
};

using namespace tmx;

extern "C" void tmx_pool_destroy(tmx_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop = true;
    }
    p->cv_job.notify_all();
    for (auto& w : p->workers)
        if (w.joinable()) w.join();
    for (auto& kv : p->jobs)
        if (kv.second->proof) tmx_proof_free(kv.second->proof);
    for (tmx_circuit* c : p->circuits) tmx_circuit_free(c);
    for (tmx_ctx* c : p->ctxs) tmx_ctx_destroy(c);
    delete p;
}

// artefact != NULL: every prover loads the `build` output (./build/main.circuit) instead of building the circuit
static int pool_create(int device, uint32_t kind, uint32_t n_max, const char* chain_id, size_t chain_id_len, uint64_t skip_max,
                       const char* artefact, unsigned in_flight, tmx_pool** out) {
    *out = nullptr;
    tmx_pool* p = new tmx_pool();
    for (unsigned k = 0; k < in_flight; k++) {
        tmx_ctx* ctx = nullptr;
        int rc = tmx_ctx_create(device, &ctx);
        if (rc) {
            tmx_pool_destroy(p);
            return rc;
        }
        p->ctxs.push_back(ctx);
        tmx_circuit* c = nullptr;
        rc = artefact ? tmx_circuit_load(ctx, artefact, &c) : tmx_circuit_build(ctx, kind, n_max, chain_id, chain_id_len, skip_max, &c);
        if (rc) {
            tmx_pool_destroy(p);
            return rc;
        }
        p->circuits.push_back(c);
    }
    for (unsigned k = 0; k < in_flight; k++) p->workers.emplace_back([p, k] { p->run(k); });
    *out = p;
    return TMX_OK;
}

extern "C" int tmx_pool_create(int device, uint32_t kind, uint32_t n_max, const char* chain_id, size_t chain_id_len,
                               uint64_t skip_max, unsigned in_flight, tmx_pool** out) {
    if (!out || !chain_id || in_flight == 0 || in_flight > 64) return fail(TMX_E_INPUT, "tmx_pool_create: bad arguments");
    return pool_create(device, kind, n_max, chain_id, chain_id_len, skip_max, nullptr, in_flight, out);
}

extern "C" int tmx_pool_create_from_artefact(int device, const char* path, unsigned in_flight, tmx_pool** out) {
    if (!out || !path || in_flight == 0 || in_flight > 64) return fail(TMX_E_INPUT, "tmx_pool_create_from_artefact: bad arguments");
    return pool_create(device, 0, 0, nullptr, 0, 0, path, in_flight, out);
}

extern "C" int tmx_pool_set_inputs(tmx_pool* p, const uint8_t* blob, size_t blob_len) {
    if (!p || !blob) return fail(TMX_E_INPUT, "tmx_pool_set_inputs: NULL argument");
    // the resident inputs are read by every proof that runs from them: replace them only while the pool is idle (nothing
    // queued AND nothing being proved), and keep the lock so that no worker can start in the middle
    std::unique_lock<std::mutex> lk(p->m);
    p->cv_done.wait(lk, [&] { return p->queue.empty() && p->in_progress == 0; });
    int first_rc = TMX_OK;
    for (tmx_circuit* c : p->circuits) {
        const int rc = tmx_circuit_set_inputs(c, blob, blob_len);
        if (rc && !first_rc) first_rc = rc;  // sizes are checked before anything is copied: all circuits fail alike
    }
    return first_rc;
}

extern "C" int tmx_pool_submit(tmx_pool* p, const uint8_t* input, size_t input_len, const uint8_t* blob, size_t blob_len,
                               uint64_t* ticket) {
    if (!p || !input || !ticket) return fail(TMX_E_INPUT, "tmx_pool_submit: NULL argument");
    auto j = std::make_shared<tmx_pool::Job>();
    j->input.assign(input, input + input_len);
    j->resident = (blob == nullptr);
    if (blob) j->blob.assign(blob, blob + blob_len);
    {
        std::lock_guard<std::mutex> lk(p->m);
        *ticket = p->next_ticket++;
        p->jobs[*ticket] = j;
        p->queue.push_back(j);
    }
    p->cv_job.notify_one();
    return TMX_OK;
}

extern "C" int tmx_pool_wait(tmx_pool* p, uint64_t ticket, tmx_proof** proof, uint8_t out32[32]) {
    if (!p || !proof || !out32) return fail(TMX_E_INPUT, "tmx_pool_wait: NULL argument");
    std::shared_ptr<tmx_pool::Job> j;
    {
        std::unique_lock<std::mutex> lk(p->m);
        auto it = p->jobs.find(ticket);
        if (it == p->jobs.end()) return fail(TMX_E_INPUT, "tmx_pool_wait: unknown ticket");
        j = it->second;
        p->cv_done.wait(lk, [&] { return j->done; });
        p->jobs.erase(it);
    }
    *proof = j->proof;
    j->proof = nullptr;
    memcpy(out32, j->out, 32);
    if (j->rc) {
        set_last_check(j->check);  // the failing gadget check, as after a direct tmx_prove on this thread
        return fail(j->rc, j->err);
    }
    return TMX_OK;
}

extern "C" unsigned tmx_pool_in_flight(const tmx_pool* p) { return p ? (unsigned)p->circuits.size() : 0; }

extern "C" uint64_t tmx_pool_launch_count(const tmx_pool* p) {
    uint64_t n = 0;
    if (p)
        for (const tmx_ctx* c : p->ctxs) n += tmx_ctx_launch_count(c);
    return n;
}

// device time of the trace commitments inside the last proof of prover k (see tmx_circuit_last_phase_ms)
extern "C" int tmx_pool_last_phase_ms(const tmx_pool* p, unsigned k, float out[6]) {
    if (!p || k >= p->circuits.size()) return fail(TMX_E_INPUT, "tmx_pool_last_phase_ms: bad arguments");
    return tmx_circuit_last_phase_ms(p->circuits[k], out);
}
