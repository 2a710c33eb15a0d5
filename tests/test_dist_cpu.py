"""CPU tests of the multi-GPU host logic: TP sharding of checkpoints (vs the oracle) and the out-of-band exchange
used to bootstrap the NVLink peer buffers, run with the gloo backend at world_size 2."""
import os
import socket

import numpy as np
import pytest

from oracle import gptq, model as omodel, ops as oops
from zhilight_b200 import dist as zdist

CFG = dict(num_layers=1, dim_model=256, num_heads=4, num_kv_heads=2, dim_head=64, dim_ff=512, vocab_size=64,
           eps=1e-5, rope_theta=10000.0, rope_llama3=None)


def _w(sd, prefix, quant):
    if quant == 5:
        qw, qz, sc, _ = gptq.to_k_major(sd[prefix + ".qweight"], sd[prefix + ".qzeros"], sd[prefix + ".scales"], None, 128)
        return gptq.dequant_k_major_f32(qw, qz, sc)
    return sd[prefix + ".weight"].astype(np.float32)


@pytest.mark.parametrize("quant", [0, 5])
@pytest.mark.parametrize("ws", [2])
def test_tp_sharding_reproduces_the_full_linear(quant, ws):
    sd = omodel.make_state_dict(CFG, quant, 128, False, seed=1)
    rng = np.random.default_rng(0)
    p = "layers.0."
    x = rng.standard_normal((3, 256)).astype(np.float32)
    shards = [zdist.shard_state_dict(sd, r, ws) for r in range(ws)]
    # column parallel: outputs concatenate
    for name in ("attn.project_q", "attn.project_k", "ff.w_in"):
        full = x @ _w(sd, p + name, quant).T
        parts = [x @ _w(shards[r], p + name, quant).T for r in range(ws)]
        np.testing.assert_allclose(np.concatenate(parts, axis=1), full, rtol=1e-6, atol=1e-6)
    # row parallel: inputs split, partial sums add up
    for name, k in (("attn.attn_out", 256), ("ff.w_out", 512)):
        xin = rng.standard_normal((3, k)).astype(np.float32)
        full = xin @ _w(sd, p + name, quant).T
        step = k // ws
        parts = [xin[:, r * step:(r + 1) * step] @ _w(shards[r], p + name, quant).T for r in range(ws)]
        np.testing.assert_allclose(sum(parts), full, rtol=1e-5, atol=1e-4)
    # replicated / vocab-parallel
    assert shards[1]["layers.0.ln_attn.weight"].shape == sd["layers.0.ln_attn.weight"].shape
    assert shards[1]["lm_head.weight"].shape[0] == CFG["vocab_size"] // ws
    np.testing.assert_array_equal(np.concatenate([s["lm_head.weight"] for s in shards]), sd["lm_head.weight"])


def test_row_parallel_needs_group_aligned_split():
    sd = omodel.make_state_dict(CFG, 5, 128, False, seed=2)
    with pytest.raises(ValueError):
        zdist.shard_tensor("layers.0.attn.attn_out.qweight", sd["layers.0.attn.attn_out.qweight"], 0, 4)   # 256/4 = 64 < g


def test_one_shot_int8_allreduce_math():
    rng = np.random.default_rng(3)
    for ws in (2, 4, 8):
        parts = [rng.standard_normal((4, 256)).astype(np.float16) for _ in range(ws)]
        a = oops.allreduce_int8_one_shot(parts)
        e = oops.allreduce_exact(parts)
        rel = np.linalg.norm(a - e) / np.linalg.norm(e)
        ref = oops.allreduce_int8_reference(parts)
        rel_ref = np.linalg.norm(ref - e) / np.linalg.norm(e)
        assert rel < 0.02 and rel <= rel_ref * 1.6, (ws, rel, rel_ref)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        # (1) the IPC-handle exchange: every rank gets every payload, in rank order
        got = zdist.exchange_bytes(bytes([rank]) * 64)
        ok1 = got == [bytes([r]) * 64 for r in range(ws)]
        # (2) the one-shot int8 protocol over the same channel: each rank contributes one quantised vector, all
        # ranks reduce in rank order and must end bit-identical and equal to the oracle
        rng = np.random.default_rng(100)
        parts = [rng.standard_normal((2, 128)).astype(np.float16) for _ in range(ws)]
        qv, sv = oops.quant_group_32(parts[rank].astype(np.float32).reshape(-1, 32))
        payloads = zdist.exchange_bytes(qv.tobytes() + sv.astype(np.float32).tobytes())
        acc = np.zeros((8, 32), np.float32)
        for pl in payloads:
            qq = np.frombuffer(pl[:256], dtype=np.int8).reshape(8, 32)
            ss = np.frombuffer(pl[256:], dtype=np.float32)
            acc = acc + qq.astype(np.float32) * ss[:, None]
        mine = oops._t(acc, "f16").reshape(2, 128)
        ok2 = np.array_equal(mine, oops.allreduce_int8_one_shot(parts))
        # (3) sharded weights from (1 process per rank) reproduce the unsharded linear after an exact all-reduce
        sd = omodel.make_state_dict(CFG, 5, 128, False, seed=4)
        sh = zdist.shard_state_dict(sd, rank, ws)
        xin = np.random.default_rng(5).standard_normal((2, 512)).astype(np.float32)
        step = 512 // ws
        part = xin[:, rank * step:(rank + 1) * step] @ _w(sh, "layers.0.ff.w_out", 5).T
        import torch
        t = torch.from_numpy(part.copy())
        dist.all_reduce(t)
        ok3 = np.allclose(t.numpy(), xin @ _w(sd, "layers.0.ff.w_out", 5).T, rtol=1e-5, atol=1e-4)
        q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_bootstrap_and_protocol():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1] and r[2] and r[3], r


def _ll_tag(step, index):
    """comm_ll_tag of zhilight_b200/csrc/comm_dev.cuh: 32-bit, never 0."""
    return ((step << 9) + index + 1) & 0xFFFFFFFF


@pytest.mark.parametrize("X", [4, 5])
@pytest.mark.parametrize("ws", [2, 4])
def test_tagged_exchange_protocol_model(ws, X):
    """Model of the exchange that rides inside the W4 GEMMs (w4a16_gemm_v3.cu tp_mode 1 / 2): every rank runs, per step,
    PUSH(x) then CONSUME(x) for x = 0 .. X-1; a word is accepted iff its tag is the tag of (step, x); consecutive exchanges
    alternate between two slots ACROSS step boundaries too (parity = (x + step * (X & 1)) & 1 -- with parity = x & 1 and an
    odd X this model finds the overwrite of an unconsumed word at the step boundary).
    Under every interleaving the dependencies allow (a rank may run ahead until it needs a word that has not arrived),
    no rank may ever read a stale word as valid and no word may be overwritten before its receiver consumed it."""
    rng = np.random.default_rng(ws)
    steps = 4
    for trial in range(200):
        inbox = [[[(0, None)] * ws for _ in range(2)] for _ in range(ws)]      # [receiver][parity][source] = (tag, payload)
        consumed = [[[True] * ws for _ in range(2)] for _ in range(ws)]       # was the current content read by its receiver?
        pc = [0] * ws                                   # program counter: 2 * (step * X + x) + (0 push | 1 consume)
        end = 2 * steps * X
        gidx = [0] * ws                                 # exchanges pushed so far, over all steps (parity = gidx of that exchange & 1)
        while any(p < end for p in pc):
            ready = []
            for r in range(ws):
                if pc[r] >= end:
                    continue
                step, x = divmod(pc[r] // 2, X)
                if pc[r] % 2 == 0:
                    ready.append(r)                     # a push never blocks
                else:
                    par = (x + step * (X & 1)) & 1
                    if all(inbox[r][par][s][0] == _ll_tag(step + 1, x) for s in range(ws)):
                        ready.append(r)
            assert ready, "deadlock"
            r = int(rng.choice(ready))
            step, x = divmod(pc[r] // 2, X)
            par = (x + step * (X & 1)) & 1
            if pc[r] % 2 == 0:
                for dst in range(ws):
                    assert consumed[dst][par][r], "overwrote a word its receiver had not consumed"
                    inbox[dst][par][r] = (_ll_tag(step + 1, x), (r, step, x))
                    consumed[dst][par][r] = False
                gidx[r] += 1
            else:
                for s in range(ws):
                    tag, payload = inbox[r][par][s]
                    assert payload == (s, step, x), "accepted a stale word"
                    consumed[r][par][s] = True
            pc[r] += 1
