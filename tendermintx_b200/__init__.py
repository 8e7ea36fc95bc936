"""tendermintx_b200 -- B200-native witness generator and Goldilocks prover for TendermintX's skip / step
circuits.  Python is a thin ctypes shim over the C ABI (include/tmx.h); torch is used only for device
memory, streams and torch.distributed plumbing.

Field elements are canonical Goldilocks u64; torch has no uint64 arithmetic so device buffers are
``torch.int64`` tensors holding the same bits (``numpy_u64.view(numpy.int64)``).
"""
import ctypes

from . import _lib
from ._lib import TmxError

__all__ = ["Context", "Circuit", "ProverPool", "InputDataFetcher", "TmxError", "lib"]

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _lib.load()
    return _LIB


def _check(rc):
    if rc != 0:
        raise TmxError(rc, lib().tmx_last_error().decode())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class Context:
    """One per GPU (mirrors `tmx_ctx`).  Kernel-level entry points take torch CUDA int64 tensors."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        self.device = device
        _check(lib().tmx_ctx_create(device, ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().tmx_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def _stream(self):
        import torch

        # torch's default stream is the legacy NULL stream; the C ABI reserves NULL for "the context's own
        # stream", so pass cudaStreamLegacy (0x1) to stay ordered with torch work.
        h = torch.cuda.current_stream(self.device).cuda_stream
        return ctypes.c_void_p(h if h else 1)

    def torch_stream(self):
        """The context's own stream as a torch ExternalStream (events recorded on it bracket tmx_prove)."""
        import torch

        return torch.cuda.ExternalStream(lib().tmx_ctx_stream(self._h), device=f"cuda:{self.device}")

    def launch_count(self):
        return int(lib().tmx_ctx_launch_count(self._h))

    def sync(self):
        _check(lib().tmx_ctx_sync(self._h))

    # ---- K1 ----
    def ntt(self, data, log_n, inverse=False):
        """In-place NTT of a [n_cols, 2^log_n] int64 CUDA tensor (each row of the tensor is one column
        of the column-major matrix), natural order in and out."""
        n_cols = data.numel() >> log_n
        _check(lib().tmx_ntt(self._h, _ptr(data), n_cols, log_n, int(inverse), self._stream()))
        return data

    def lde(self, values, log_n, rate_bits, out=None, coeffs=None):
        import torch

        n_cols = values.numel() >> log_n
        if out is None:
            out = torch.empty((n_cols, 1 << (log_n + rate_bits)), dtype=torch.int64, device=values.device)
        _check(lib().tmx_lde(self._h, _ptr(values), _ptr(out), _ptr(coeffs), n_cols, log_n, rate_bits, self._stream()))
        return out

    # ---- K2 ----
    def poseidon_merkle(self, cols, log_rows, cap_height, digests=None):
        import torch

        n_cols = cols.numel() >> log_rows
        cnt = lib().tmx_merkle_digest_count(log_rows, cap_height)
        if digests is None:
            digests = torch.empty((cnt, 4), dtype=torch.int64, device=cols.device)
        _check(lib().tmx_poseidon_merkle(self._h, _ptr(cols), n_cols, log_rows, cap_height, _ptr(digests), self._stream()))
        return digests

    def poseidon_permute(self, states):
        _check(lib().tmx_poseidon_permute(self._h, _ptr(states), states.numel() // 12, self._stream()))
        return states

    # ---- K6 / K7 / K8 ----
    @staticmethod
    def trace_dims(kind, n_max):
        d = (ctypes.c_size_t * 6)()
        _check(lib().tmx_trace_dims(kind, n_max, d))
        return [(d[0], d[1]), (d[2], d[3]), (d[4], d[5])]

    def witness_generate(self, blob, kind, n_max):
        """blob: bytes (host).  Returns (sha256, sha512, ed25519 tables as [cols, rows] int64 CUDA tensors, aux
        uint8 CUDA tensor)."""
        import torch

        dev = f"cuda:{self.device}"
        d_blob = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        tabs = [torch.empty((c, r), dtype=torch.int64, device=dev) for r, c in self.trace_dims(kind, n_max)]
        aux = torch.zeros(lib().tmx_witness_aux_bytes(n_max), dtype=torch.uint8, device=dev)
        _check(lib().tmx_witness_generate(self._h, _ptr(d_blob), kind, n_max, _ptr(tabs[0]), _ptr(tabs[1]),
                                          _ptr(tabs[2]), _ptr(aux), self._stream()))
        return tabs, aux


    def sha512_trace(self, blob, kind, n_max):
        """K7 alone: the SHA-512 table as a [cols, rows] int64 CUDA tensor."""
        import torch

        dev = f"cuda:{self.device}"
        d_blob = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        r, c = self.trace_dims(kind, n_max)[1]
        tab = torch.empty((c, r), dtype=torch.int64, device=dev)
        _check(lib().tmx_sha512_trace(self._h, _ptr(d_blob), kind, n_max, _ptr(tab), self._stream()))
        return tab

    # ---- K3 (fold) ----
    def fri_fold(self, values, log_cosets, shift, beta):
        """One arity-16 FRI fold: values [16 << log_cosets, 2] int64 CUDA tensor (extension elements, bit-reversed order) on
        shift * <w> -> [1 << log_cosets, 2] on shift^16 * <w^16>."""
        import torch

        out = torch.empty((1 << log_cosets, 2), dtype=torch.int64, device=values.device)
        b = (ctypes.c_uint64 * 2)(*[int(x) for x in beta])
        _check(lib().tmx_fri_fold(self._h, _ptr(values), log_cosets, int(shift), b, _ptr(out), self._stream()))
        return out


KIND_STEP, KIND_SKIP = 0, 1
T_SHA256, T_SHA512, T_ED, T_LOGIC, T_RANGE = range(5)
SKIP_MAX = 100800  # REF circuits/config.rs:12
CHECK_NAMES = ["OK", "SKIP_DISTANCE", "TRUSTED_HEADER_PROOF", "TRUSTED_VALHASH", "TRUSTED_THRESHOLD", "SIGNATURE",
               "VALHASH", "VALHASH_PROOF", "THRESHOLD", "SIGN_BYTES", "CHAIN_ID", "HEIGHT", "LAST_BLOCK_ID",
               "NEXT_VALHASH", "VOTING_OVERFLOW", "VARINT_SIGN", "ROUND_SIGN", "INPUT"]


class TendermintConfig:
    """REF circuits/config.rs:3-32."""

    def __init__(self, chain_id, skip_max=SKIP_MAX):
        self.chain_id = chain_id.encode() if isinstance(chain_id, str) else bytes(chain_id)
        self.skip_max = skip_max


CelestiaConfig = TendermintConfig("celestia")
Mocha4Config = TendermintConfig("mocha-4")


def verify_proof(kind, n_max, config, proof, public_input, output):
    """CPU verification from bare circuit parameters (no GPU, no context).  Raises TmxError if rejected."""
    proof, public_input, output = bytes(proof), bytes(public_input), bytes(output)
    _check(lib().tmx_verify_params(kind, n_max, config.chain_id, len(config.chain_id), config.skip_max, proof, len(proof),
                                   public_input, len(public_input), output))


def logic_trace(kind, n_max, config, public_input, blob, force=False):
    """The logic table (every plain-gate gadget of verify_skip / verify_step, one instance per row) of one proof, filled on
    the host exactly as tmx_prove does.  Returns (cells as a flat u64 list in [column][row] order, status): status is 0 or
    the id of the first failing check; with `force` the table is completed past it."""
    public_input, blob = bytes(public_input), bytes(blob)
    st = ctypes.c_int(0)
    args = (kind, n_max, config.chain_id, len(config.chain_id), public_input, blob, int(force))
    cells = lib().tmx_logic_trace(*args, None, 0, ctypes.byref(st))
    if cells == 0:
        raise TmxError(lib().tmx_last_error().decode() or "this circuit has no logic table")
    out = (ctypes.c_uint64 * cells)()
    lib().tmx_logic_trace(*args, out, cells, ctypes.byref(st))
    return out, st.value


def blob_size(kind, n_max):
    return 920 + 240 * n_max + (48 * n_max if kind == KIND_SKIP else 0)


class InputDataFetcher:
    """Fixture mode of the reference's InputDataFetcher [REF circuits/input/mod.rs:37-43,96-116]: `fixture_path`
    holds <height>/commit.json and <height>/validators_<page>.json.  Methods return the packed off-chain blob."""

    def __init__(self, fixture_path):
        self.fixture_path = str(fixture_path).encode()

    def header_hash(self, block):
        out = (ctypes.c_uint8 * 32)()
        _check(lib().tmx_header_hash_from_fixture(self.fixture_path, block, out))
        return bytes(out)

    def get_skip_inputs(self, n_max, trusted_block, trusted_hash, target_block):
        buf = (ctypes.c_uint8 * blob_size(KIND_SKIP, n_max))()
        _check(lib().tmx_skip_inputs_from_fixture(self.fixture_path, n_max, trusted_block, bytes(trusted_hash), target_block,
                                                  buf, len(buf)))
        return bytes(buf)

    def is_valid_skip(self, start_block, target_block):
        """REF circuits/input/tendermint_utils.rs:444-482."""
        v = ctypes.c_int()
        _check(lib().tmx_is_valid_skip_from_fixture(self.fixture_path, start_block, target_block, ctypes.byref(v)))
        return bool(v.value)

    def find_block_to_request(self, start_block, max_end_block):
        """REF circuits/input/mod.rs:158-186: highest block to request a skip to (start_block + 1 = request a step)."""
        b = ctypes.c_uint64()
        _check(lib().tmx_find_block_to_request(self.fixture_path, start_block, max_end_block, ctypes.byref(b)))
        return int(b.value)

    def get_step_inputs(self, n_max, prev_block, prev_hash):
        buf = (ctypes.c_uint8 * blob_size(KIND_STEP, n_max))()
        _check(lib().tmx_step_inputs_from_fixture(self.fixture_path, n_max, prev_block, bytes(prev_hash), buf, len(buf)))
        return bytes(buf)


class Circuit:
    """`SkipCircuit::<N, C>` / `StepCircuit::<N, C>` [REF circuits/skip.rs:103-143, circuits/step.rs:90-127]:
    build() -> prove(input, offchain blob) -> verify(proof, input, output)."""

    def __init__(self, ctx, kind, n_max, config):
        self.ctx, self.kind, self.n_max, self.config = ctx, kind, n_max, config
        self._h = ctypes.c_void_p()
        _check(lib().tmx_circuit_build(ctx.handle, kind, n_max, config.chain_id, len(config.chain_id), config.skip_max,
                                       ctypes.byref(self._h)))

    @classmethod
    def build(cls, ctx, kind, n_max, config):
        return cls(ctx, kind, n_max, config)

    def close(self):
        if self._h:
            lib().tmx_circuit_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def digest(self):
        out = (ctypes.c_uint64 * 4)()
        _check(lib().tmx_circuit_digest(self._h, out))
        return list(out)

    def save(self, path):
        _check(lib().tmx_circuit_save(self._h, path.encode()))

    @classmethod
    def load(cls, ctx, path, kind, n_max, config):
        """`<bin> prove` side of the CLI contract: load ./build/main.circuit written by `build` [REF succinct.json:8-9]."""
        h = ctypes.c_void_p()
        _check(lib().tmx_circuit_load(ctx.handle, str(path).encode(), ctypes.byref(h)))
        c = cls.__new__(cls)
        c.ctx, c.kind, c.n_max, c.config, c._h = ctx, kind, n_max, config, h
        return c

    def set_inputs(self, blob):
        """Upload the off-chain inputs once; prove(public_input, None) then runs from the HBM-resident copy."""
        blob = bytes(blob)
        _check(lib().tmx_circuit_set_inputs(self._h, blob, len(blob)))

    def last_phase_ms(self):
        """Device time of the trace commitments inside the last prove(): [(lde_ms, merkle_ms)] per table."""
        out = (ctypes.c_float * 6)()
        _check(lib().tmx_circuit_last_phase_ms(self._h, out))
        return [(out[2 * t], out[2 * t + 1]) for t in range(3)]

    def prove(self, public_input, blob):
        """Returns (proof bytes, output header bytes).  Raises TmxError(TMX_E_UNSAT) with `.check` set to the name
        of the failing gadget check where the reference's witness generation would panic."""
        p = ctypes.c_void_p()
        out = (ctypes.c_uint8 * 32)()
        public_input = bytes(public_input)
        blob = bytes(blob) if blob is not None else None
        rc = lib().tmx_prove(self._h, public_input, len(public_input), blob, len(blob) if blob is not None else 0,
                             ctypes.byref(p), out)
        if rc != 0:
            err = TmxError(rc, lib().tmx_last_error().decode())
            err.check = CHECK_NAMES[lib().tmx_last_check()]
            raise err
        n = lib().tmx_proof_size(p)
        buf = (ctypes.c_uint8 * n)()
        _check(lib().tmx_proof_bytes(p, buf, n))
        lib().tmx_proof_free(p)
        return bytes(buf), bytes(out)

    def prove_fixture(self, public_input, fixture_path):
        """`prove input.json` with off-chain inputs read from a fixture directory (the async-hint path)."""
        p = ctypes.c_void_p()
        out = (ctypes.c_uint8 * 32)()
        public_input = bytes(public_input)
        rc = lib().tmx_prove_fixture(self._h, public_input, len(public_input), str(fixture_path).encode(), ctypes.byref(p), out)
        if rc != 0:
            err = TmxError(rc, lib().tmx_last_error().decode())
            err.check = CHECK_NAMES[lib().tmx_last_check()]
            raise err
        n = lib().tmx_proof_size(p)
        buf = (ctypes.c_uint8 * n)()
        _check(lib().tmx_proof_bytes(p, buf, n))
        lib().tmx_proof_free(p)
        return bytes(buf), bytes(out)

    def verify(self, proof, public_input, output):
        proof, public_input, output = bytes(proof), bytes(public_input), bytes(output)
        _check(lib().tmx_verify(self._h, proof, len(proof), public_input, len(public_input), output))

    # ---- kernel-level entry points that need the circuit's constant / periodic columns ----
    def table_shape(self, table):
        """(rows, first-round columns, constant columns, second-round columns); rows = 0 when the table is absent."""
        out = (ctypes.c_size_t * 4)()
        _check(lib().tmx_circuit_table_shape(self._h, table, out))
        return tuple(out)

    def bus_count(self, table, trace):
        """Histogram (int32 CUDA tensor, 2^16 + 2^11 + 2^8 + 2 entries: the 16-, 11-, 8- and 1-bit tables) of the range lookups of a first-round trace, and the
        out-of-range flag."""
        import torch

        hist = torch.zeros((1 << 16) + (1 << 11) + (1 << 8) + 2, dtype=torch.int32, device=trace.device)
        bad = ctypes.c_int(0)
        _check(lib().tmx_bus_count(self._h, table, _ptr(trace), _ptr(hist), ctypes.byref(bad), self.ctx._stream()))
        return hist, bool(bad.value)

    def bus_aux(self, table, trace, beta, gamma):
        """Second-round trace ([aux cols, rows] int64 CUDA tensor) and the table's bus total (two ints)."""
        import torch

        rows, _, _, a = self.table_shape(table)
        aux = torch.empty((a, rows), dtype=torch.int64, device=trace.device)
        b = (ctypes.c_uint64 * 2)(*[int(x) for x in beta])
        g = (ctypes.c_uint64 * 2)(*[int(x) for x in gamma])
        tot = (ctypes.c_uint64 * 2)()
        _check(lib().tmx_bus_aux(self._h, table, _ptr(trace), b, g, _ptr(aux), tot, self.ctx._stream()))
        return aux, (tot[0], tot[1])

    def quotient(self, table, lde_main, lde_aux, total, beta, gamma, alpha):
        """K5: quotient values on the LDE coset, [2, 2 rows] int64 CUDA tensor in natural order."""
        import torch

        rows = self.table_shape(table)[0]
        out = torch.empty((2, 2 * rows), dtype=torch.int64, device=lde_main.device)
        arr = lambda v: (ctypes.c_uint64 * 2)(*[int(x) for x in v])
        _check(lib().tmx_quotient(self._h, table, _ptr(lde_main), _ptr(lde_aux), arr(total), arr(beta), arr(gamma), arr(alpha),
                                  _ptr(out), self.ctx._stream()))
        return out


class ProverPool:
    """Several provers in flight on ONE GPU (`tmx_pool_*`, native: one tmx_ctx + circuit + host thread per prover, one
    FIFO queue).

    A proof is a chain of kernels interleaved with Fiat-Shamir round trips to the host (cap -> challenge -> next
    kernel), so a single prover leaves the GPU idle for about a sixth of the proof.  Proofs are independent
    (SURVEY section 8e), so a service keeps `in_flight` of them going: the gaps of one are filled by the kernels of the
    others.  Nothing is shared between the provers and no proof changes.
    """

    def __init__(self, device, kind, n_max, config, in_flight=4, artefact=None):
        self.device, self.kind, self.n_max, self.config = device, kind, n_max, config
        self._h = ctypes.c_void_p()
        if artefact is None:
            _check(lib().tmx_pool_create(device, kind, n_max, config.chain_id, len(config.chain_id), config.skip_max, in_flight,
                                         ctypes.byref(self._h)))
        else:
            _check(lib().tmx_pool_create_from_artefact(device, str(artefact).encode(), in_flight, ctypes.byref(self._h)))
        self.in_flight = int(lib().tmx_pool_in_flight(self._h))

    def launch_count(self):
        return int(lib().tmx_pool_launch_count(self._h))

    def last_phase_ms(self, prover=0):
        """Device time of the trace commitments inside prover `prover`'s last proof: [(lde_ms, merkle_ms)] per table."""
        out = (ctypes.c_float * 6)()
        _check(lib().tmx_pool_last_phase_ms(self._h, prover, out))
        return [(out[2 * t], out[2 * t + 1]) for t in range(3)]

    def set_inputs(self, blob):
        blob = bytes(blob)
        _check(lib().tmx_pool_set_inputs(self._h, blob, len(blob)))

    def submit(self, public_input, blob):
        """Queue one proof (blob None = the resident inputs).  Returns a ticket for wait()."""
        t = ctypes.c_uint64()
        public_input = bytes(public_input)
        blob = bytes(blob) if blob is not None else None
        _check(lib().tmx_pool_submit(self._h, public_input, len(public_input), blob, len(blob) if blob is not None else 0,
                                     ctypes.byref(t)))
        return int(t.value)

    def wait(self, ticket):
        """Returns (proof bytes, output header bytes) of a submitted proof; raises TmxError as Circuit.prove would."""
        p = ctypes.c_void_p()
        out = (ctypes.c_uint8 * 32)()
        rc = lib().tmx_pool_wait(self._h, ticket, ctypes.byref(p), out)
        if rc != 0:
            err = TmxError(rc, lib().tmx_last_error().decode())
            err.check = CHECK_NAMES[lib().tmx_last_check()]
            raise err
        n = lib().tmx_proof_size(p)
        buf = (ctypes.c_uint8 * n)()
        _check(lib().tmx_proof_bytes(p, buf, n))
        lib().tmx_proof_free(p)
        return bytes(buf), bytes(out)

    def prove_many(self, statements):
        """statements: list of (public_input, blob or None).  Returns the list of (proof, output) in order; the first
        failing statement (e.g. TMX_E_UNSAT) raises after every queued proof has finished."""
        tickets = [self.submit(p, b) for p, b in statements]
        results, first_error = [], None
        for t in tickets:
            try:
                results.append(self.wait(t))
            except TmxError as e:
                results.append(None)
                first_error = first_error or e
        if first_error is not None:
            raise first_error
        return results

    def close(self):
        if self._h:
            lib().tmx_pool_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
