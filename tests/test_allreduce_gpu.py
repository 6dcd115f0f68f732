"""msda_allreduce_f32 (csrc/allreduce.cu): the gradient all-reduce over NVLink peer memory against NCCL's all_reduce on the same
data -- both algorithms (two-shot P2P, multimem through the switch when the box has multicast), sub-ranges, sum and mean, ragged
sizes, many back-to-back calls (the flag words must return to zero), and inside a replayed CUDA graph.

Needs two GPUs: skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_allreduce_gpu.py -m gpu`
(log committed under profiles/).  The CPU suite checks that the library exports the entry points (tests/test_abi.py)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    from mdqe_cvpr2023_b200.collectives import PeerAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    log = []
    try:
        n = 1_000_004
        ar = PeerAllReduce(n, dev, algo="p2p", n_ctas=4)
        algos = ["p2p"] + (["multimem"] if ar._mc else [])
        log.append(f"multicast {'yes' if ar._mc else 'no'}")
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        for algo in algos:
            ar.algo = algo
            for off, cnt, mean, ctas in ((0, None, True, 4), (0, None, False, 1), (4096, 40, False, 3), (12, 999_000, True, 16), (0, 4, False, 8)):
                x = torch.randn(ar.numel, device=dev, generator=g)
                ar.buffer.copy_(x)
                want = x.clone()
                cnt_ = ar.numel - off if cnt is None else cnt
                seg = want[off:off + cnt_].clone()
                dist.all_reduce(seg)
                want[off:off + cnt_] = seg / world if mean else seg
                torch.cuda.synchronize()
                dist.barrier()
                ar.all_reduce_(off, cnt, mean=mean, n_ctas=ctas)
                torch.cuda.synchronize()
                ar.check()
                err = float((ar.buffer - want).abs().max())
                assert err <= 1e-5, f"{algo} off={off} cnt={cnt} mean={mean}: max abs err {err}"
                dist.barrier()
            # 50 back-to-back reductions, then the same inside a replayed graph: sum of ones doubles every time
            ar.buffer.fill_(1.0)
            torch.cuda.synchronize(); dist.barrier()
            for _ in range(10):
                ar.all_reduce_(mean=False)
            torch.cuda.synchronize(); ar.check()
            assert float(ar.buffer.min()) == float(ar.buffer.max()) == float(world) ** 10, f"{algo}: chained sums {float(ar.buffer.max())}"
            ar.buffer.fill_(1.0)
            torch.cuda.synchronize(); dist.barrier()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ar.all_reduce_(mean=True)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                ar.all_reduce_(mean=False)
                ar.all_reduce_(mean=True)
            for _ in range(5):
                gr.replay()
            torch.cuda.synchronize(); ar.check()
            # warm-up mean of ones = 1; every replay sums (x world) and then averages identical values (x 1)
            assert float(ar.buffer.min()) == float(ar.buffer.max()) == float(world) ** 5, f"{algo}: graph replays {float(ar.buffer.max())}"
            log.append(f"{algo} ok")
            dist.barrier()
        if rank == 0:
            ret["log"] = log
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_peer_allreduce_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)          # every GPU of the box: the P2P kernel is instantiated per world size
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        if p.is_alive():
            p.kill()
            pytest.fail("rank did not finish")
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    print("peer all-reduce:", "; ".join(ret["log"]))
