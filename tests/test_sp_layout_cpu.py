"""CPU, world_size 2 over gloo: the sharding the library reports (flux2b_sp_layout, the same function the forward uses)
tiles the joint sequence exactly, and an all-to-all of the [dest][token][q|k|v][heads/P*128] layout written by the QKV
epilogue yields, on every rank, all tokens for its own heads."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S_txt, S_img, H, q):
    sys.path.insert(0, os.path.join(ROOT, "flux-2-swift-mlx_b200"))
    import flux2b
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = flux2b.sp_layout(world, rank, S_txt, S_img, H)
        Hp, Dp, D = L.heads_per_rank, L.heads_per_rank * 128, H * 128
        # the full QKV every rank would compute for ALL tokens (deterministic), in the reference's [txt | img] order
        g = torch.Generator().manual_seed(0)
        full = torch.randn(S_txt + S_img, 3 * D, generator=g)
        rows = list(range(L.txt_row0, L.txt_row0 + L.txt_rows)) + [S_txt + r for r in range(L.img_row0, L.img_row0 + L.img_rows)]
        local = full[rows]                                              # [local_rows, 3D], rows = [txt shard | img shard]
        # what EPI_QKV_ROPE stores under sp: [dest][token][which][Hp*128]
        send = local.reshape(L.local_rows, 3, world, Dp).permute(2, 0, 1, 3).contiguous()
        assert send[0].numel() == L.qkv_chunk_elems
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        # rank-major token order [txt_0 | img_0 | txt_1 | img_1 ...], my heads only
        order = []
        for r in range(world):
            Lr = flux2b.sp_layout(world, r, S_txt, S_img, H)
            order += list(range(Lr.txt_row0, Lr.txt_row0 + Lr.txt_rows)) + [S_txt + i for i in range(Lr.img_row0, Lr.img_row0 + Lr.img_rows)]
        want = full[order].reshape(S_txt + S_img, 3, world, Dp)[:, :, rank]
        ok = torch.equal(recv.reshape(S_txt + S_img, 3, Dp), want) and sorted(order) == list(range(S_txt + S_img))
        # second exchange: O [S, Dp] (rank-major rows) -> my tokens, all heads in columns [j*Dp, (j+1)*Dp)
        o_mine = want[:, 0].contiguous()                                 # stand-in for the attention output of my heads
        orecv = torch.empty(world, L.local_rows, Dp)
        dist.all_to_all_single(orecv, o_mine.reshape(world, L.local_rows, Dp).contiguous())
        assert orecv[0].numel() == L.o_chunk_elems
        cat = orecv.permute(1, 0, 2).reshape(L.local_rows, D)
        ok = ok and torch.equal(cat, local[:, :D])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("S_txt,S_img,H", [(512, 4096, 24), (64, 256, 8)])
def test_ulysses_layout_roundtrip_world2(S_txt, S_img, H):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, S_txt, S_img, H, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_sp_layout_rejects_bad_shapes():
    sys.path.insert(0, os.path.join(ROOT, "flux-2-swift-mlx_b200"))
    import flux2b
    for args in ((3, 0, 512, 4096, 24), (2, 0, 511, 4096, 24), (8, 0, 512, 4096, 20), (2, 2, 512, 4096, 24), (9, 0, 576, 4608, 72)):
        with pytest.raises(flux2b.Flux2Error) as e:
            flux2b.sp_layout(*args)
        assert e.value.case == "invalidConfiguration"
    L = flux2b.sp_layout(8, 3, 512, 16384, 48)  # Dev @2048^2 on 8 GPUs
    assert (L.txt_row0, L.txt_rows, L.img_row0, L.img_rows, L.heads_per_rank) == (192, 64, 6144, 2048, 6)
    assert L.qkv_chunk_elems == 2112 * 3 * 768 and L.o_chunk_elems == 2112 * 768
