"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise -- run with `gpurun --gpus 2`):
the NVLink one-shot all-reduce / all-gather against exact sums, and a TP=2 decode against the single-GPU oracle."""
import os
import socket

import numpy as np
import pytest

from oracle import model as omodel, ops as oops
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu

TINY = dict(num_layers=2, dim_model=256, num_heads=4, num_kv_heads=2, dim_head=64, dim_ff=512, vocab_size=512,
            eps=1e-5, rope_theta=10000.0, rope_llama3=None)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    res = {}
    try:
        from zhilight_b200 import dist as zdist
        from zhilight_b200.llama import LlamaDecoder
        dev = torch.device("cuda", rank)
        comm = zdist.TPComm(4 * 8192, rank, ws)
        rng = np.random.default_rng(0)
        parts = [rng.standard_normal((4, 8192)).astype(np.float16) for _ in range(ws)]
        resid = rng.standard_normal((4, 8192)).astype(np.float16)
        mine = torch.from_numpy(parts[rank]).to(dev)
        # fp16 payload: rank-ordered fp32 sum, rounded once -> equals the exact oracle bit for bit
        for it in range(5):      # several epochs: both parities of the double-buffered inbox
            out = comm.allreduce(mine, None)
            torch.cuda.synchronize()
        res["ar_exact"] = bool(np.array_equal(out.float().cpu().numpy(), oops.allreduce_exact(parts)))
        out = comm.allreduce(mine, torch.from_numpy(resid).to(dev))
        exp = oops.residual_add(oops.allreduce_exact(parts), resid)
        res["ar_resid"] = bool(np.array_equal(out.float().cpu().numpy(), exp))
        out8 = comm.allreduce(mine, None, int8=True)
        res["ar_int8"] = bool(np.array_equal(out8.float().cpu().numpy(), oops.allreduce_int8_one_shot(parts)))
        g = comm.allgather(torch.full((8,), rank, dtype=torch.int32, device=dev))
        res["ag"] = bool((g.cpu().numpy() == np.arange(ws)[:, None]).all())
        # NCCL agrees (the reference's reduce_sum path)
        t = mine.clone()
        dist.all_reduce(t)
        res["vs_nccl"] = float(rel_l2(out.float().cpu().numpy() - resid.astype(np.float32), t.float().cpu().numpy()))

        # TP decode vs the single-GPU oracle.  W4 decode rides the exchange inside the GEMMs (partial tiles pushed into the
        # peers' inboxes from the o_proj / w_out epilogues, reduced in the next GEMM's staging); ZL_TP_UNFUSED=1 keeps the
        # stand-alone one-shot kernels: same arithmetic in the same order, so the two must agree bit for bit.
        fused_logits = {}
        for quant, unfused in ((5, False), (5, True), (0, False)):
            if unfused:
                os.environ["ZL_TP_UNFUSED"] = "1"
            else:
                os.environ.pop("ZL_TP_UNFUSED", None)
            sd = omodel.make_state_dict(TINY, quant, 128, False, seed=2)
            dec = LlamaDecoder(quant_type=quant, max_batch=2, max_seq=32, tp_rank=rank, tp_size=ws, **TINY)
            c2 = zdist.TPComm(2 * TINY["dim_model"], rank, ws)
            dec.set_comm(c2)
            dec.load_state_dict(zdist.shard_state_dict(sd, rank, ws))
            orc = omodel.OracleLlama(TINY, sd, quant, 128, False, "f16", fuse_norm=True)
            tok = np.array([5, 99], dtype=np.int32)
            worst = 0.0
            agree = True
            same = True
            for step in range(6):
                pos = np.full(2, step, dtype=np.int32)
                nxt, logits = dec.decode(tok, pos, want_logits=True)       # logits: this rank's vocab shard
                ref = orc.decode(tok, pos)
                vs = TINY["vocab_size"] // ws
                worst = max(worst, rel_l2(logits, ref[:, rank * vs:(rank + 1) * vs]))
                order = np.argsort(ref, axis=1)
                for b in range(2):
                    if ref[b, order[b, -1]] - ref[b, order[b, -2]] > 2e-2:
                        agree = agree and nxt[b] == order[b, -1]
                if quant == 5 and not unfused:
                    fused_logits[step] = logits.copy()
                if quant == 5 and unfused:
                    same = same and np.array_equal(logits, fused_logits[step])
                tok = nxt
            kern = dec.stats(2)[1]
            if quant == 5 and unfused:
                res["tp_fused_equals_unfused"] = bool(same)
                res["kernels_unfused"] = int(kern)
            else:
                res["tp_quant%d" % quant] = (float(worst), bool(agree))
                if quant == 5:
                    res["kernels_fused"] = int(kern)
            dec.close()
            c2.close()
        os.environ.pop("ZL_TP_UNFUSED", None)
        # TP chunked prefill: with tp_size > 1 the two-half / two-stream pipeline is on by default --
        # the reduce stream all-reduces one half's partial sums while the compute stream runs the other half
        sd = omodel.make_state_dict(TINY, 5, 128, False, seed=4)
        dec = LlamaDecoder(quant_type=5, max_batch=1, max_seq=96, tp_rank=rank, tp_size=ws, prefill_chunk=32, **TINY)
        c3 = zdist.TPComm(32 * TINY["dim_model"], rank, ws)
        dec.set_comm(c3)
        dec.load_state_dict(zdist.shard_state_dict(sd, rank, ws))
        orc = omodel.OracleLlama(TINY, sd, 5, 128, False, "f16", fuse_norm=True)
        prompt = np.random.default_rng(8).integers(0, TINY["vocab_size"], size=52).astype(np.int32)
        nxt, logits = dec.prefill(0, prompt, want_logits=True)
        for p_, t_ in enumerate(prompt):
            ref = orc.decode(np.array([t_]), [p_])
        vs = TINY["vocab_size"] // ws
        res["tp_prefill"] = float(rel_l2(logits, ref[:, rank * vs:(rank + 1) * vs]))
        nxt2, logits2 = dec.decode(np.array([int(np.argmax(ref[0]))], np.int32), np.array([len(prompt)], np.int32),
                                   want_logits=True)
        ref2 = orc.decode(np.array([int(np.argmax(ref[0]))]), [len(prompt)])
        res["tp_prefill_then_decode"] = float(rel_l2(logits2, ref2[:, rank * vs:(rank + 1) * vs]))
        dec.close()
        c3.close()
        comm.close()
    except Exception as e:   # pragma: no cover
        import traceback
        res["error"] = traceback.format_exc()
    finally:
        q.put((rank, res))
        dist.destroy_process_group()


@pytest.mark.parametrize("ws", [2])
def test_tp_exchange_and_decode(lib, cuda, ws):
    import torch
    if torch.cuda.device_count() < ws:
        pytest.skip("needs %d GPUs" % ws)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(ws):
        res = results[r]
        assert "error" not in res, res.get("error")
        assert res["ar_exact"] and res["ar_resid"] and res["ar_int8"] and res["ag"], res
        assert res["vs_nccl"] < 2e-3, res
        assert res["tp_quant5"][0] < 3e-3 and res["tp_quant5"][1], res
        assert res["tp_quant0"][0] < 3e-3 and res["tp_quant0"][1], res
        assert res["tp_fused_equals_unfused"], res
        assert res["kernels_fused"] < res["kernels_unfused"], res        # 2 * layers - 1 exchange kernels are gone
        assert res["tp_prefill"] < 5e-3 and res["tp_prefill_then_decode"] < 5e-3, res
