"""world_size-2 gloo test of the N>1 host logic (no GPU): artefact broadcast, proof sharding, max-over-ranks."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, hashlib
    sys.path.insert(0, __ROOT__)
    import torch.distributed as dist
    from tendermintx_b200 import sharding
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    artefact = bytes(range(256)) * 37 if rank == 0 else b""
    got = sharding.broadcast_bytes(artefact, 0)
    assert hashlib.sha256(got).hexdigest() == hashlib.sha256(bytes(range(256)) * 37).hexdigest()
    mine = sharding.assign_proofs(8, rank, world)
    assert mine == list(range(rank, 8, world))
    t = sharding.max_over_ranks([10.0 + rank, 5.0 - rank])
    assert t == [10.0 + world - 1, 5.0]
    sys.stdout.write("rank %d ok %d %s\\n" % (rank, len(got), mine)); sys.stdout.flush()
    dist.destroy_process_group()
""").replace("__ROOT__", repr(ROOT))


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29613", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_assignment_covers_every_proof_once():
    from tendermintx_b200 import sharding

    for world in (1, 2, 4, 8):
        seen = sorted(j for r in range(world) for j in sharding.assign_proofs(11, r, world))
        assert seen == list(range(11))
