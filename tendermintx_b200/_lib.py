"""ctypes loader for libtmx.so (the C ABI declared in include/tmx.h).

There is no CPU fallback: if the shared library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C tendermintx_b200/csrc``) importing this module
raises, and without a CUDA device ``tmx_ctx_create`` fails.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtmx.so")

TMX_OK, TMX_E_INPUT, TMX_E_UNSAT, TMX_E_CUDA, TMX_E_IO, TMX_E_VERIFY = range(6)
_STATUS_NAMES = {
    TMX_E_INPUT: "TMX_E_INPUT",
    TMX_E_UNSAT: "TMX_E_UNSAT",
    TMX_E_CUDA: "TMX_E_CUDA",
    TMX_E_IO: "TMX_E_IO",
    TMX_E_VERIFY: "TMX_E_VERIFY",
}


class TmxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (make -C tendermintx_b200/csrc). "
            "tendermintx_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, u64p, sz, u32, i32 = c.c_void_p, c.c_void_p, c.c_size_t, c.c_uint, c.c_int
    sig = {
        "tmx_last_error": (c.c_char_p, []),
        "tmx_version": (c.c_char_p, []),
        "tmx_ctx_create": (i32, [i32, c.POINTER(vp)]),
        "tmx_ctx_destroy": (None, [vp]),
        "tmx_ctx_sync": (i32, [vp]),
        "tmx_ctx_stream": (vp, [vp]),
        "tmx_ctx_launch_count": (c.c_uint64, [vp]),
        "tmx_ntt": (i32, [vp, u64p, sz, u32, i32, vp]),
        "tmx_lde": (i32, [vp, u64p, u64p, u64p, sz, u32, u32, vp]),
        "tmx_merkle_digest_count": (sz, [u32, u32]),
        "tmx_poseidon_merkle": (i32, [vp, u64p, sz, u32, u32, u64p, vp]),
        "tmx_poseidon_permute": (i32, [vp, u64p, sz, vp]),
        "tmx_host_poseidon_permute": (i32, [u64p, sz, i32]),
        "tmx_trace_dims": (i32, [u32, u32, c.POINTER(sz)]),
        "tmx_witness_aux_bytes": (sz, [u32]),
        "tmx_sha256_trace": (i32, [vp, vp, u32, u32, u64p, vp, vp]),
        "tmx_ed25519_trace": (i32, [vp, vp, u32, u32, u64p, u64p, vp, vp]),
        "tmx_witness_generate": (i32, [vp, vp, u32, u32, u64p, u64p, u64p, vp, vp]),
        "tmx_sha512_trace": (i32, [vp, vp, u32, u32, u64p, vp]),
        "tmx_quotient": (i32, [vp, i32, u64p, u64p, vp, vp, vp, vp, u64p, vp]),
        "tmx_bus_aux": (i32, [vp, i32, u64p, vp, vp, u64p, vp, vp]),
        "tmx_bus_count": (i32, [vp, i32, u64p, vp, c.POINTER(c.c_int), vp]),
        "tmx_fri_fold": (i32, [vp, u64p, u32, c.c_uint64, vp, u64p, vp]),
        "tmx_circuit_table_shape": (i32, [vp, i32, c.POINTER(sz)]),
        "tmx_circuit_artefact": (sz, [u32, u32, c.c_char_p, sz, c.c_uint64, vp, sz]),
        "tmx_logic_trace": (sz, [u32, u32, c.c_char_p, sz, c.c_char_p, c.c_char_p, i32, vp, sz, c.POINTER(c.c_int)]),
        "tmx_pow_grind": (i32, [vp, vp, i32, u32, c.POINTER(c.c_uint64), vp]),
        "tmx_circuit_build": (i32, [vp, u32, u32, c.c_char_p, sz, c.c_uint64, c.POINTER(vp)]),
        "tmx_circuit_free": (None, [vp]),
        "tmx_circuit_digest": (i32, [vp, vp]),
        "tmx_circuit_save": (i32, [vp, c.c_char_p]),
        "tmx_circuit_load": (i32, [vp, c.c_char_p, c.POINTER(vp)]),
        "tmx_prove": (i32, [vp, vp, sz, vp, sz, c.POINTER(vp), vp]),
        "tmx_last_check": (i32, []),
        "tmx_circuit_last_phase_ms": (i32, [vp, vp]),
        "tmx_circuit_set_inputs": (i32, [vp, vp, sz]),
        "tmx_header_hash_from_fixture": (i32, [c.c_char_p, c.c_uint64, vp]),
        "tmx_skip_inputs_from_fixture": (i32, [c.c_char_p, u32, c.c_uint64, vp, c.c_uint64, vp, sz]),
        "tmx_step_inputs_from_fixture": (i32, [c.c_char_p, u32, c.c_uint64, vp, vp, sz]),
        "tmx_is_valid_skip_from_fixture": (i32, [c.c_char_p, c.c_uint64, c.c_uint64, c.POINTER(c.c_int)]),
        "tmx_find_block_to_request": (i32, [c.c_char_p, c.c_uint64, c.c_uint64, c.POINTER(c.c_uint64)]),
        "tmx_pool_create": (i32, [i32, u32, u32, c.c_char_p, sz, c.c_uint64, u32, c.POINTER(vp)]),
        "tmx_pool_create_from_artefact": (i32, [i32, c.c_char_p, u32, c.POINTER(vp)]),
        "tmx_pool_destroy": (None, [vp]),
        "tmx_pool_set_inputs": (i32, [vp, vp, sz]),
        "tmx_pool_submit": (i32, [vp, vp, sz, vp, sz, c.POINTER(c.c_uint64)]),
        "tmx_pool_wait": (i32, [vp, c.c_uint64, c.POINTER(vp), vp]),
        "tmx_pool_in_flight": (u32, [vp]),
        "tmx_pool_launch_count": (c.c_uint64, [vp]),
        "tmx_pool_last_phase_ms": (i32, [vp, u32, vp]),
        "tmx_prove_fixture": (i32, [vp, vp, sz, c.c_char_p, c.POINTER(vp), vp]),
        "tmx_proof_size": (sz, [vp]),
        "tmx_proof_bytes": (i32, [vp, vp, sz]),
        "tmx_proof_free": (None, [vp]),
        "tmx_verify": (i32, [vp, vp, sz, vp, sz, vp]),
        "tmx_verify_params": (i32, [u32, u32, c.c_char_p, sz, c.c_uint64, vp, sz, vp, sz, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


EXPORTED_SYMBOLS = [
    "tmx_last_error", "tmx_version", "tmx_ctx_create", "tmx_ctx_destroy", "tmx_ctx_sync",
    "tmx_ctx_stream", "tmx_ctx_launch_count", "tmx_ntt", "tmx_lde", "tmx_merkle_digest_count", "tmx_poseidon_merkle",
    "tmx_poseidon_permute", "tmx_host_poseidon_permute", "tmx_trace_dims", "tmx_witness_aux_bytes", "tmx_sha256_trace", "tmx_ed25519_trace",
    "tmx_witness_generate", "tmx_sha512_trace", "tmx_quotient", "tmx_bus_aux", "tmx_bus_count", "tmx_fri_fold", "tmx_circuit_table_shape", "tmx_circuit_artefact", "tmx_logic_trace", "tmx_pow_grind", "tmx_circuit_build", "tmx_circuit_free", "tmx_circuit_digest",
    "tmx_circuit_save", "tmx_circuit_load", "tmx_prove", "tmx_last_check", "tmx_circuit_last_phase_ms", "tmx_circuit_set_inputs", "tmx_header_hash_from_fixture", "tmx_skip_inputs_from_fixture",
    "tmx_step_inputs_from_fixture", "tmx_is_valid_skip_from_fixture", "tmx_find_block_to_request", "tmx_prove_fixture",
    "tmx_pool_create", "tmx_pool_create_from_artefact", "tmx_pool_destroy", "tmx_pool_set_inputs", "tmx_pool_submit", "tmx_pool_wait", "tmx_pool_in_flight",
    "tmx_pool_launch_count", "tmx_pool_last_phase_ms", "tmx_proof_size", "tmx_proof_bytes",
    "tmx_proof_free", "tmx_verify", "tmx_verify_params",
]
