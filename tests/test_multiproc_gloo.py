"""CPU-only, world_size 2 over gloo: the multi-GPU host path (tile partition -> all_gather -> compose; chunk-range meshing
-> count all_gather -> variable-length gather) with the oracle standing in for the kernels.  Checks that the N-rank result
equals the 1-rank result byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenes

W, H = 160, 90


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import orc
    from mesoengine_b200 import partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    origin, dims, params = scenes.sphere_scene(256)
    vol = orc.Volume(origin, dims).voxelize(orc.SDF_SPHERE, params, granularity=orc.GRAN_VOXEL, nthreads=2)  # replicated volume
    eyes, ctr = scenes.orbit_eyes(origin, dims, 8)
    cam = orc.camera_uniform(eyes[3], ctr, width=W, height=H)
    rs = orc.ray_setup(cam, origin, W, H)
    # --- raymarch: this rank renders only its tiles, packs them, all ranks gather, rank 0 composes ---
    frame = np.zeros((H, W), dtype=orc.HitRecord)
    for t in partition.rank_tiles(W, H, rank, world):
        x0, y0, x1, y1 = partition.tile_rect(W, H, int(t))
        frame[y0:y1, x0:x1] = vol.raymarch(rs, W, H, rect=(x0, y0, x1, y1), nthreads=1)[y0:y1, x0:x1]
    packed = partition.pack_tiles(frame.view(np.uint32).reshape(H, W, 4), rank, world)
    mine = torch.from_numpy(packed.astype(np.int64))
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    composed = partition.compose_tiles(np.stack([g.numpy().astype(np.uint32) for g in gathered]), W, H)
    # --- meshing: chunk c belongs to rank c % world; counts all-gathered, lists gathered at prefix offsets ---
    occ = vol.occ()
    keys = []
    for c in partition.rank_chunks(vol.nchunks, rank, world):
        bits = np.unpackbits(occ[c].view(np.uint8), bitorder="little")
        keys += [int(c) * 4096 + int(b) for b in np.nonzero(bits)[0]]
    mine_chunks = np.array(list(partition.rank_chunks(vol.nchunks, rank, world)), dtype=np.int64)
    quads = np.concatenate([vol.mesh_bricks(np.array(keys, dtype=np.uint64)), vol.mesh_chunk_faces(mine_chunks)]).view(np.uint32).reshape(-1, 4)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(quads)], dtype=torch.int64))
    counts = [int(c.item()) for c in counts]
    cap = max(counts)
    buf = torch.zeros((cap, 4), dtype=torch.int64); buf[: len(quads)] = torch.from_numpy(quads.astype(np.int64))
    allq = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(allq, buf)
    merged = np.concatenate([a.numpy()[:n] for a, n in zip(allq, counts)]).astype(np.uint32)
    if rank == 0:
        full = vol.raymarch(rs, W, H, nthreads=2).view(np.uint32).reshape(H, W, 4)
        ref_q = orc.sort_quads(vol.mesh(nthreads=2))
        got_q = orc.sort_quads(np.ascontiguousarray(merged).view(orc.Quad).reshape(-1))
        q.put((bool(np.array_equal(composed, full)), bool(got_q.tobytes() == ref_q.tobytes()), len(ref_q)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_equal_one_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    frames_equal, quads_equal, nq = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert frames_equal and quads_equal and nq > 0
