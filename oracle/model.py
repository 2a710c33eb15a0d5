"""CPU oracle: a whole Llama decode step (numpy), assembled from oracle.gptq / oracle.ops.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows the reference's single-stream decode order
(src/nn/block/block.cpp:86-143, src/nn/attention/attention.cpp:846-964,
src/nn/feedforward/feedforward.cpp:113-137, src/model/llama.cpp:75-165): every operator output is
rounded to the activation dtype T exactly where the reference materialises a T tensor.
"""
import numpy as np

from . import gptq, ops

F32 = np.float32


def _f32(a, dtype):
    """Checkpoint tensor -> fp32 values.  bf16 tensors travel as uint16 bit patterns (numpy has no bf16)."""
    a = np.asarray(a)
    if a.dtype == np.uint16:
        return (a.astype(np.uint32) << 16).view(np.float32)
    return a.astype(F32)


def to_bf16_bits(x):
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


class OracleLlama:
    def __init__(self, cfg, state_dict, quant_type=0, group_size=128, sym=False, dtype="f16", fuse_norm=False):
        """cfg: dict with num_layers, dim_model, num_heads, num_kv_heads, dim_head, dim_ff, vocab_size, eps,
        rope_theta, rope_llama3.  state_dict: HF/ZhiLight-named numpy tensors."""
        self.c = dict(cfg)
        self.dtype = dtype
        self.sd = state_dict
        self.quant_type = quant_type
        self.group_size = group_size
        self.sym = sym
        # fuse_norm=True restates OUR fused kernel's rounding points (x*w_ln rounded to T feeds the GEMM, the
        # fp32 accumulator is scaled by rsqrt(mean(x^2)+eps)); False is the reference's operator order.
        self.fuse_norm = fuse_norm and quant_type in (5, 6)
        self.w = {}
        self.kv = {}          # task -> list over layers of (k (cap,Hkv,d), v)

    def _weight(self, prefix):
        """(N, K) fp32 dequantized weight of one Linear."""
        if prefix in self.w:
            return self.w[prefix]
        sd = self.sd
        if self.quant_type == 5:
            qw, qz, sc, _ = gptq.to_k_major(sd[prefix + ".qweight"], sd[prefix + ".qzeros"], sd[prefix + ".scales"],
                                            None, self.group_size)
            w = gptq.dequant_k_major_f32(qw, qz, sc, self.sym)
        elif self.quant_type == 6:
            qw, qz, sc, _ = gptq.to_k_major(sd[prefix + ".qweight"], sd[prefix + ".qzeros"], sd[prefix + ".scales"],
                                            None, self.group_size, is_awq=True)
            w = gptq.dequant_k_major_f32(qw, qz, sc, False)
        else:
            w = _f32(sd[prefix + ".weight"], self.dtype)
        self.w[prefix] = w
        return w

    def _linear(self, x, prefix):
        if self.quant_type == 2:
            # AutoInt8 (linear.cpp:521-550, 560-636): per-row int8 weights made at load, scale kept in T; per-token
            # int8 activations; s32 GEMM; scale-back; bias added in T
            if prefix not in self.w:
                wq, ws = ops.int8_quant_per_token(_f32(self.sd[prefix + ".weight"], self.dtype))
                self.w[prefix] = (wq, ops._t(ws, self.dtype))
            wq, ws = self.w[prefix]
            b = self.sd.get(prefix + ".bias")
            return ops.int8_linear(x, wq, ws, self.dtype, None if b is None else _f32(b, self.dtype))
        if self.quant_type == 7:
            # Fp8Linear (linear.cpp:1660-1695): per-tensor dynamic e4m3 activations, e4m3 weights + scalar scale
            if prefix not in self.w:
                self.w[prefix] = ops.e4m3_decode(self.sd[prefix + ".weight"])
            b = self.sd.get(prefix + ".bias")
            return ops.fp8_linear(x, self.w[prefix], np.float32(self.sd[prefix + ".weight_scale"]).reshape(-1)[0],
                                  self.dtype, None if b is None else _f32(b, self.dtype))
        y = np.asarray(x, F32) @ self._weight(prefix).T
        b = self.sd.get(prefix + ".bias")
        if b is not None:
            y = y + _f32(b, self.dtype)[None, :]
        return ops._t(y, self.dtype)

    def _norm_linears(self, h, ln_w, prefixes):
        """[T(W_i . rmsnorm(h))] for several Linears sharing one RMSNorm'd input."""
        c, T = self.c, self.dtype
        if not self.fuse_norm:
            xn = ops.rmsnorm(h, ln_w, c["eps"], 1.0, T)
            return [self._linear(xn, p) for p in prefixes]
        h = np.asarray(h, F32)
        xw = ops._t(h * np.asarray(ln_w, F32), T)
        rstd = F32(1.0) / np.sqrt((h * h).sum(-1, keepdims=True, dtype=F32) / F32(h.shape[-1]) + F32(c["eps"]))
        outs = []
        for p in prefixes:
            y = (xw @ self._weight(p).T) * rstd
            b = self.sd.get(p + ".bias")
            if b is not None:
                y = y + _f32(b, T)[None, :]
            outs.append(ops._t(y, T))
        return outs

    def decode(self, tokens, positions, tasks=None):
        """One step for B tasks; returns logits (B, V) fp32.  Task b appends its K/V at positions[b]."""
        c, T = self.c, self.dtype
        b = len(tokens)
        tasks = list(range(b)) if tasks is None else tasks
        d, hq, hkv = c["dim_head"], c["num_heads"], c["num_kv_heads"]
        emb = _f32(self.sd["token_embedding.weight"], T)
        h = emb[np.asarray(tokens)]
        cos, sin = ops.rope_cos_sin(np.asarray(positions), d, c["rope_theta"], c.get("rope_llama3"))
        scale = F32(1.0 / np.sqrt(d))
        for l in range(c["num_layers"]):
            p = "layers.%d." % l
            q, k, v = self._norm_linears(h, _f32(self.sd[p + "ln_attn.weight"], T),
                                         [p + "attn.project_q", p + "attn.project_k", p + "attn.project_v"])
            qkv = np.concatenate([q, k, v], axis=1)
            q, k, v = ops.split_qkv_rope(qkv, cos, sin, hq, hkv, d, True, T)
            ao = np.zeros((b, hq * d), F32)
            for i, task in enumerate(tasks):
                kvs = self.kv.setdefault(task, [None] * c["num_layers"])
                if kvs[l] is None:
                    kvs[l] = (np.zeros((0, hkv, d), F32), np.zeros((0, hkv, d), F32))
                kb, vb = kvs[l]
                pos = int(positions[i])
                if kb.shape[0] <= pos:
                    pad = np.zeros((pos + 1 - kb.shape[0], hkv, d), F32)
                    kb, vb = np.concatenate([kb, pad]), np.concatenate([vb, pad])
                kb[pos] = k[i].reshape(hkv, d)
                vb[pos] = v[i].reshape(hkv, d)
                kvs[l] = (kb, vb)
                lb = pos + 1
                o = ops.decode_attention(q[i].reshape(1, 1, hq, d), [kb], [vb], [lb],
                                         [np.ones((1, lb), np.int8)], scale, hq // hkv, T)
                ao[i] = o.reshape(-1)
            o = self._linear(ao, p + "attn.attn_out")
            h = ops.residual_add(h, o, T)
            g, u = self._norm_linears(h, _f32(self.sd[p + "ln_ff.weight"], T), [p + "ff.w_in", p + "ff.w_gated"])
            act = ops.silu_mul(g, u, T)
            dn = self._linear(act, p + "ff.w_out")
            h = ops.residual_add(h, dn, T)
        xn = ops.rmsnorm(h, _f32(self.sd["output_layernorm.weight"], T), c["eps"], 1.0, T)
        lm = self.sd.get("lm_head.weight", self.sd["token_embedding.weight"])
        return np.asarray(xn, F32) @ _f32(lm, T).T


def make_state_dict(cfg, quant_type=0, group_size=128, sym=False, seed=0, tied=False, dtype="f16", scale_range=None):
    """Random checkpoint with ZhiLight names (zhilight/loader.py:250-358).  scale_range = (lo, hi) of the GPTQ group scales
    (wide layers need small ones to keep fp16 activations finite)."""
    rng = np.random.default_rng(seed)
    c = cfg
    d_model, d = c["dim_model"], c["dim_head"]
    sd = {}

    def cast(a):
        return to_bf16_bits(a) if dtype == "bf16" else np.asarray(a).astype(np.float16)

    def dense(n, k):
        return cast(rng.standard_normal((n, k)) * 0.05)

    def linear(prefix, k, n, s):
        if quant_type == 5:
            extra = {} if scale_range is None else dict(scale_lo=scale_range[0], scale_hi=scale_range[1])
            qw, qz, sc, gi = gptq.make_gptq_checkpoint(k, n, group_size, sym, s, **extra)
            sd[prefix + ".qweight"], sd[prefix + ".qzeros"], sd[prefix + ".scales"] = qw, qz, sc
        elif quant_type == 6:
            r = np.random.default_rng(s)
            sd[prefix + ".qweight"] = r.integers(0, 2 ** 32, size=(k, n // 8), dtype=np.uint64).astype(np.uint32).view(np.int32)
            sd[prefix + ".qzeros"] = r.integers(0, 2 ** 32, size=(k // group_size, n // 8), dtype=np.uint64).astype(np.uint32).view(np.int32)
            sd[prefix + ".scales"] = (0.005 + 0.015 * r.random((k // group_size, n))).astype(np.float16)
        elif quant_type == 7:
            r = np.random.default_rng(s)
            w8 = r.integers(0, 256, size=(n, k)).astype(np.uint8) & 0xBF     # exponent MSB cleared: |w| < 2, no NaN codes
            sd[prefix + ".weight"] = w8
            sd[prefix + ".weight_scale"] = np.array([0.02 + 0.02 * r.random()], dtype=np.float32)
        else:
            sd[prefix + ".weight"] = dense(n, k)

    s = seed * 1000
    for l in range(c["num_layers"]):
        p = "layers.%d." % l
        sd[p + "ln_attn.weight"] = cast(1.0 + 0.1 * rng.standard_normal(d_model))
        sd[p + "ln_ff.weight"] = cast(1.0 + 0.1 * rng.standard_normal(d_model))
        linear(p + "attn.project_q", d_model, c["num_heads"] * d, s + 1)
        linear(p + "attn.project_k", d_model, c["num_kv_heads"] * d, s + 2)
        linear(p + "attn.project_v", d_model, c["num_kv_heads"] * d, s + 3)
        linear(p + "attn.attn_out", c["num_heads"] * d, d_model, s + 4)
        linear(p + "ff.w_in", d_model, c["dim_ff"], s + 5)
        linear(p + "ff.w_gated", d_model, c["dim_ff"], s + 6)
        linear(p + "ff.w_out", c["dim_ff"], d_model, s + 7)
        s += 10
    sd["token_embedding.weight"] = dense(c["vocab_size"], d_model)
    sd["output_layernorm.weight"] = cast(1.0 + 0.1 * rng.standard_normal(d_model))
    if not tied:
        sd["lm_head.weight"] = dense(c["vocab_size"], d_model)
    return sd
